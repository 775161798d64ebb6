#!/bin/bash
# Round-2 closing session: the driver's own commands (suite, smoke, both bench arms), the other BASELINE configurations, ncu launch
# lists and full captures of the kernels of the SHIPPED build (-> profiles/r02_*).
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/final_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/final_smoke.txt
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2> gpurun_out/final_err.log | tee gpurun_out/final_bench_reference.json | cut -c1-200
python bench.py --gpus 1 2>> gpurun_out/final_err.log | tee gpurun_out/final_bench.json | cut -c1-300
python bench.py --holes 0 --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>> gpurun_out/final_err.log | tee gpurun_out/final_bench_noholes.json | cut -c1-200
for w in chain2048 batched512 footprint4096 footprint4096_offset0 footprint_polygon4096 slope8192; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>> gpurun_out/final_err.log | tee gpurun_out/final_bench_$w.json | cut -c1-200
done
python bench.py --workload plugin_chain --steps 5 --warmup 2 2>> gpurun_out/final_err.log | tee gpurun_out/final_bench_plugin_chain.json | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 24 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 40 --csv --log-file gpurun_out/final_fp_launches.csv \
    python bench.py --workload footprint_polygon4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 40 --csv --log-file gpurun_out/final_fpc_launches.csv \
    python bench.py --workload footprint4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_chain_fused -s 2 -c 1 -f -o gpurun_out/final_fused \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fixup_t2 -s 2 -c 1 -f -o gpurun_out/final_t2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_fast -s 1 -c 1 -f -o gpurun_out/final_sweep \
    python bench.py --workload footprint4096 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_poly_tile -s 2 -c 1 -f -o gpurun_out/final_poly \
    python bench.py --workload footprint_polygon4096 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f7.log 2>&1
tail -3 gpurun_out/final_err.log
