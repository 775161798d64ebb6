#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/err1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('holes', round(d['value']), d['roofline']['kernel_ms'], d['roofline']['fixup_kernel_ms'], d['config']['slow_path_cells_per_launch'])"
python bench.py --steps 20 --warmup 3 --holes 0 --no-cpu-baseline --no-e2e 2>> gpurun_out/err1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('noholes', round(d['value']), d['roofline']['kernel_ms'], d['roofline']['fixup_kernel_ms'], d['config']['slow_path_cells_per_launch'])"
python tools/dev_scale_check.py 2>&1 | tail -6
tail -3 gpurun_out/err1.log
