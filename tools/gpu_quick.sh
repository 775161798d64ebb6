#!/bin/bash
mkdir -p gpurun_out
python tools/dev_check_fused.py 2>&1 | grep -E "tier|slope|rough" | cut -c1-200
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2> gpurun_out/err1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('holes', d['value'], d['roofline']['kernel_ms'], d['roofline']['fixup_kernel_ms'], d['config']['slow_path_cells_per_launch'])"
python bench.py --steps 20 --warmup 3 --holes 0 --no-cpu-baseline --no-e2e 2>> gpurun_out/err1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('noholes', d['value'], d['roofline']['kernel_ms'], d['roofline']['fixup_kernel_ms'], d['config']['slow_path_cells_per_launch'])"
tail -3 gpurun_out/err1.log
if [ "$1" == "ncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:k_chain_fused -s 2 -c 1 -f -o gpurun_out/prof_fused python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
fi
if [ "$2" != "" ]; then
make -C traversability_estimation_b200/csrc -B TE_WPC=$2 -s 2>&1 | grep -E "error"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/err1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('wpc$2 holes', d['value'], d['roofline']['kernel_ms'], d['roofline']['fixup_kernel_ms'])"
fi
