#!/bin/bash
mkdir -p gpurun_out
python tools/dev_check_fused.py 2>&1 | grep -E "tier|slope|step|rough|trav'|'nz'" | cut -c1-200
python bench.py --steps 10 --warmup 3 --no-e2e 2> gpurun_out/err1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('holes', d['value'], d['roofline']['kernel_ms'], d['roofline']['fixup_kernel_ms'], d['config']['slow_path_cells_per_launch'], d['cpu_baseline'])"
python bench.py --steps 10 --warmup 3 --holes 0 --no-cpu-baseline --no-e2e 2>> gpurun_out/err1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('noholes', d['value'], d['roofline']['kernel_ms'], d['roofline']['fixup_kernel_ms'], d['config']['slow_path_cells_per_launch'])"
tail -3 gpurun_out/err1.log
