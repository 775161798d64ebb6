#!/bin/bash
# Round-2 session 6: footprint after the shared-memory ring walk / column masks.
mkdir -p gpurun_out
python -m pytest tests/test_footprint_gpu.py tests/test_parity_scale_gpu.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/tests6.txt
for w in footprint4096 footprint4096_offset0; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/err6.log | tee gpurun_out/bench6_$w.json | cut -c1-260
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 30 --csv --log-file gpurun_out/fp6_launches.csv \
    python bench.py --workload footprint4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fp6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_tile -s 1 -c 1 -f -o gpurun_out/prof_sweep_tile \
    python bench.py --workload footprint4096 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fp6b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pred_heavy -s 1 -c 1 -f -o gpurun_out/prof_pred_heavy \
    python bench.py --workload footprint4096 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fp6c.log 2>&1
tail -3 gpurun_out/err6.log
