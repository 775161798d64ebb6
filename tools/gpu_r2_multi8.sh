#!/bin/bash
# The driver's launch at N = 8 (strong scaling of the fixed 8192^2 map), once as the driver runs it and once with the halo inline.
mkdir -p gpurun_out
N=${1:-8}
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29810 bench.py --gpus $N --steps 400 --warmup 10 2> gpurun_out/err_n8.log | tee gpurun_out/bench_drv_n8.json | cut -c1-300
tail -2 gpurun_out/err_n8.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus $N --steps 200 --warmup 10 --no-e2e --halo-overlap 0 2> gpurun_out/err_n8b.log | tee gpurun_out/bench_n8_inline.json | cut -c1-300
tail -2 gpurun_out/err_n8b.log | cut -c1-300
