#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for N in "$@"; do
  if [ "$N" == "1" ]; then
    python bench.py --gpus 1 --steps 20 --warmup 3 2> gpurun_out/err_b1.log > gpurun_out/bench_n1.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 2> gpurun_out/err_b$N.log > gpurun_out/bench_n$N.json
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 3 --scaling strong --no-e2e 2>> gpurun_out/err_b$N.log > gpurun_out/bench_strong_n$N.json
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --impl reference --steps 1 --warmup 0 2>> gpurun_out/err_b$N.log | tail -1 | cut -c1-120
  fi
  tail -2 gpurun_out/err_b$N.log | cut -c1-200
  python - <<PY
import json
for f in ("bench_n$N","bench_strong_n$N"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "Mcells/s", round(d["ms_per_step"],3), "ms", "e2e", (d.get("e2e") or {}).get("value"), d["config"]["workload"][:60])
    except Exception as e: print(f, "ERR", e)
PY
done
