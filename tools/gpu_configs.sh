#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for w in chain2048 batched512 footprint4096; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2> gpurun_out/err_$w.log | tee gpurun_out/bench_$w.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['value']), 'Mcells/s', round(d['ms_per_step'],3), 'ms', 'frac', round(d['roofline']['frac'],3), d['config'].get('slow_path_cells_per_launch'))"
  tail -2 gpurun_out/err_$w.log
done
