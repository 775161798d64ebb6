#!/bin/bash
# One GPU session: parity tests, bench lines, ncu launch list + full capture of the dominant kernel.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_err.log | tee gpurun_out/bench_holes.json
python bench.py --steps 10 --warmup 3 --holes 0 --no-cpu-baseline --no-e2e 2>> gpurun_out/bench_err.log | tee gpurun_out/bench_noholes.json
tail -5 gpurun_out/bench_err.log
if [ "$1" == "ncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_chain_fused -s 2 -c 1 -f -o gpurun_out/prof_fused \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
fi
