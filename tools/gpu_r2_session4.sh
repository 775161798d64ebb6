#!/bin/bash
# Round-2 session 4: full GPU suite (tiled footprint sweep, inclination check, small-launch plans), bench lines of the other
# BASELINE configurations, footprint launch list, two more work-queue plans.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/tests4.txt
for w in footprint4096 footprint4096_offset0 chain2048 batched512 slope8192; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/err4.log | tee gpurun_out/bench4_$w.json | cut -c1-260
done
python bench.py --workload plugin_chain --steps 5 --warmup 2 2>> gpurun_out/err4.log | tee gpurun_out/bench4_plugin_chain.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 30 --csv --log-file gpurun_out/fp4_launches.csv \
    python bench.py --workload footprint4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fp4.log 2>&1
for segs in "" "512:0.75,40:0.2,16:0.05" "512:0.75,24:0.2,16:0.05" "496:0.75,40:0.2,16:0.05"; do
  TE_FUSED_SEGS="$segs" TE_B200_LIBRARY=$PWD/traversability_estimation_b200/libte_b200_calib.so timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/err4.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('calib', '[$segs]', round(d['value']), d['ms_per_step'], d['roofline'].get('kernel_ms'), d['roofline'].get('fixup_kernel_ms'))" | tee -a gpurun_out/segs4.txt
done
tail -3 gpurun_out/err4.log
