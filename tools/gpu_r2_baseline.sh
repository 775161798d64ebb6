#!/bin/bash
# Round-2 first session: evidence on the shipped 12-warp build (bench line, launch list, ncu full of the three chain kernels).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1
python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_err.log | tee gpurun_out/r2_base_bench.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 24 --csv --log-file gpurun_out/r2_base_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_chain_fused -s 2 -c 1 -f -o gpurun_out/r2_base_fused \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
tail -3 gpurun_out/bench_err.log
