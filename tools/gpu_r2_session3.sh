#!/bin/bash
# Round-2 session 3: pipe-rate micro-benchmark, segment plans of the work queue (TE_FUSED_SEGS is read by calibration builds only ->
# here through the variant built with -DTE_CALIBRATION), footprint launch list.
mkdir -p gpurun_out
timeout 120 tools/experiments/pipe_rates | tee gpurun_out/pipe_rates.txt
for lib in calib straightcalib; do
  for segs in "" "504:0.8,40:0.15,16:0.05" "248:0.8,40:0.15,16:0.05" "160:0.8,48:0.15,16:0.05" "120:0.8,40:0.15,16:0.05" "504:0.9,32:0.07,16:0.03" "320:0.85,40:0.1,16:0.05"; do
    TE_FUSED_SEGS="$segs" TE_B200_LIBRARY=$PWD/traversability_estimation_b200/libte_b200_$lib.so timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e 2>> gpurun_out/err3.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', '[$segs]', round(d['value']), d['ms_per_step'], d['roofline'].get('kernel_ms'), d['roofline'].get('fixup_kernel_ms'))" | tee -a gpurun_out/segs.txt
  done
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 30 --csv --log-file gpurun_out/fp_launches.csv \
    python bench.py --workload footprint4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fp.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 30 --csv --log-file gpurun_out/b512_launches.csv \
    python bench.py --workload batched512 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b512.log 2>&1
tail -3 gpurun_out/err3.log
