#!/bin/bash
# Multi-GPU validation of the driver's launch at N = $1 (strong scaling, overlapped peer-mapped halo) + inline/NCCL variants.
mkdir -p gpurun_out
N=$1
for extra in "" "--halo-overlap 0" "--halo nccl"; do
  tag=$(echo "$extra" | tr -d ' -')
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 200 --warmup 10 --no-e2e $extra 2> gpurun_out/err_mm$N$tag.log > gpurun_out/bench_mm${N}_$tag.json
  tail -2 gpurun_out/err_mm$N$tag.log | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_mm${N}_$tag.json").read().strip().splitlines()[-1])
    print("N=$N [$extra]", round(d["value"]), "Mcells/s", round(d["ms_per_step"],4), "ms halo_ms", d.get("halo_ms"), "fused", d["roofline"]["kernel_ms"], "fix", d["roofline"]["fixup_kernel_ms"])
except Exception as e: print("N=$N [$extra] ERR", e)
PY
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N --steps 400 --warmup 10 2> gpurun_out/err_drv2_$N.log | tee gpurun_out/bench_drv2_n$N.json | cut -c1-300
tail -2 gpurun_out/err_drv2_$N.log | cut -c1-300
python -m pytest tests/test_halo_gpu.py tests/test_slab_batched_gpu.py -x -q -m gpu 2>&1 | tail -2
