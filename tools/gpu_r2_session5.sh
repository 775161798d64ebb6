#!/bin/bash
# Round-2 session 5: suite + bench lines after the PDL / counter-block / tiled-sweep changes.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/tests5.txt
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>> gpurun_out/err5.log | tee gpurun_out/bench5_chain8192.json | cut -c1-200
for w in footprint4096 footprint4096_offset0 chain2048 batched512; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/err5.log | tee gpurun_out/bench5_$w.json | cut -c1-260
done
python bench.py --rows 8192 --cols 1024 --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>> gpurun_out/err5.log | tee gpurun_out/bench5_slab1024.json | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 30 --csv --log-file gpurun_out/fp5_launches.csv \
    python bench.py --workload footprint4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fp5.log 2>&1
python tools/dev_scale_check.py 2>&1 | tail -6 | tee gpurun_out/scale5.txt
tail -3 gpurun_out/err5.log
