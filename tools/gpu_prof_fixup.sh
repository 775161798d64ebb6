#!/bin/bash
# ncu captures of the fix-up tiers (and a launch list) on the default bench map.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 24 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fixup_t2 -s 2 -c 1 -f -o gpurun_out/r2_t2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_t2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fixup_cells -s 2 -c 1 -f -o gpurun_out/r2_t3 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_t3.log 2>&1
tail -2 gpurun_out/ncu_t3.log
