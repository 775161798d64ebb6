#!/bin/bash
# The driver's GPU checks: full suite, smoke, one default bench line.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/suite_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/suite_smoke.txt
python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline 2> gpurun_out/suite_err.log | tee gpurun_out/suite_bench.json | cut -c1-250
tail -2 gpurun_out/suite_err.log
