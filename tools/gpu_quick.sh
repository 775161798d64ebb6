#!/bin/bash
mkdir -p gpurun_out
export TE_FUSED_SPLIT=1
timeout 300 python tools/dev_check_fused.py 2>&1 | grep -E "tier|slope|rough|Error|error" | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/err1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('split holes', round(d['value']), d['roofline']['kernel_ms'], d['roofline']['fixup_kernel_ms'], d['config']['slow_path_cells_per_launch'])"
timeout 300 python bench.py --steps 20 --warmup 3 --holes 0 --no-cpu-baseline --no-e2e 2>> gpurun_out/err1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('split noholes', round(d['value']), d['roofline']['kernel_ms'], d['roofline']['fixup_kernel_ms'])"
tail -3 gpurun_out/err1.log
unset TE_FUSED_SPLIT
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/err1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('single holes', round(d['value']), d['roofline']['kernel_ms'], d['roofline']['fixup_kernel_ms'])"
