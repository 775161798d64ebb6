#!/bin/bash
# Multi-GPU session: driver-style launches of bench.py at N = $@ (strong scaling is the default), peer-mapped halo vs NCCL.
mkdir -p gpurun_out
for N in "$@"; do
  for halo in ipc nccl; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 200 --warmup 10 --halo $halo --no-e2e 2> gpurun_out/err_m$N$halo.log > gpurun_out/bench_m${N}_$halo.json
    tail -2 gpurun_out/err_m$N$halo.log | cut -c1-300
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_m${N}_$halo.json").read().strip().splitlines()[-1])
    print("N=$N $halo", round(d["value"]), "Mcells/s", round(d["ms_per_step"],4), "ms halo_ms", d.get("halo_ms"), d["config"]["tiling"][:90])
except Exception as e: print("N=$N $halo ERR", e)
PY
  done
done
# the driver's exact command at N = last (with e2e)
N=${@: -1}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus $N --steps 400 --warmup 10 2> gpurun_out/err_drv$N.log | tee gpurun_out/bench_drv_n$N.json | cut -c1-400
tail -2 gpurun_out/err_drv$N.log | cut -c1-300
