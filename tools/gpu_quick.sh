#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/err1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('holes', d['value'], d['roofline']['kernel_ms'], d['roofline']['fixup_kernel_ms'], d['config']['slow_path_cells_per_launch'], 'e2e', d['e2e']['value'])"
tail -3 gpurun_out/err1.log
