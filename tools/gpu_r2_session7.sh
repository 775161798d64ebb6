#!/bin/bash
# Round-2 session 7: polygon sweep tests, global-memory vs tiled circular sweep (calibration build), launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_footprint_gpu.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/tests7.txt
for w in footprint4096 footprint4096_offset0; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/err7.log | tee gpurun_out/bench7_$w.json | cut -c1-230
  TE_FOOTPRINT_TILE=1 TE_B200_LIBRARY=$PWD/traversability_estimation_b200/libte_b200_calib.so python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/err7.log | tee gpurun_out/bench7_tile_$w.json | cut -c1-230
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 30 --csv --log-file gpurun_out/fp7_launches.csv \
    python bench.py --workload footprint4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fp7.log 2>&1
tail -3 gpurun_out/err7.log
