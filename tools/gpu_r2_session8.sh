#!/bin/bash
# Round-2 session 8: footprint kernels after the addressing / run-compression changes.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_footprint_gpu.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/tests8.txt
for w in footprint4096 footprint4096_offset0 footprint_polygon4096; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/err8.log | tee gpurun_out/bench8_$w.json | cut -c1-230
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 40 --csv --log-file gpurun_out/fp8_launches.csv \
    python bench.py --workload footprint4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fp8.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 40 --csv --log-file gpurun_out/fp8p_launches.csv \
    python bench.py --workload footprint_polygon4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fp8p.log 2>&1
tail -3 gpurun_out/err8.log
