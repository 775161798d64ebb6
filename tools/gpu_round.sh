#!/bin/bash
# One GPU session: parity tests, bench lines, ncu launch list + full capture of the dominant kernel.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 2> gpurun_out/bench_err.log | tee gpurun_out/bench_holes.json | cut -c1-300
python bench.py --steps 50 --warmup 5 --holes 0 --no-cpu-baseline --no-e2e 2>> gpurun_out/bench_err.log | tee gpurun_out/bench_noholes.json | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench_err.log | tee gpurun_out/bench_reference.json | cut -c1-300
for w in chain2048 batched512 footprint4096 slope8192; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/bench_err.log | tee gpurun_out/bench_$w.json | cut -c1-200
done
tail -5 gpurun_out/bench_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 24 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_chain_fused -s 2 -c 1 -f -o gpurun_out/prof_fused \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:k_fixup_t2 -s 2 -c 1 -f -o gpurun_out/prof_t2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_t2.log 2>&1
tail -2 gpurun_out/ncu_full.log
