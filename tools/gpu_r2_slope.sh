#!/bin/bash
# One short session: full suite on the build with the constant-bank acos coefficients in k_slope_stream, A/B bench of te_slope
# against the previous library (TE_B200_LIBRARY), one ncu capture of the new kernel.  Every leg under its own timeout.
mkdir -p gpurun_out
timeout 85 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/s_tests.txt; cat gpurun_out/s_tests.txt
B="python bench.py --workload slope8192 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e"
timeout 25 $B > gpurun_out/s_new.json 2> gpurun_out/s_err.log; cut -c1-200 gpurun_out/s_new.json
TE_B200_LIBRARY=$PWD/traversability_estimation_b200/libte_b200_prev.so timeout 25 $B > gpurun_out/s_prev.json 2>> gpurun_out/s_err.log; cut -c1-200 gpurun_out/s_prev.json
timeout 35 ncu --set full --clock-control none --import-source on -k regex:k_slope_stream -s 1 -c 1 -f -o gpurun_out/slope_stream \
  python bench.py --workload slope8192 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/s_ncu.log 2>&1
tail -2 gpurun_out/s_ncu.log
