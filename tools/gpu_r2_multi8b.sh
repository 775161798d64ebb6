#!/bin/bash
# N = 8 with the closing build: the driver's launch (strong scaling of the fixed 8192^2 map) and BASELINE config 4 (256 x 512^2
# maps sharded by map, no communication).
mkdir -p gpurun_out
N=${1:-8}
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29820 bench.py --gpus $N --steps 400 --warmup 10 2> gpurun_out/err_n8c.log | tee gpurun_out/bench_final_n8.json | cut -c1-300
tail -2 gpurun_out/err_n8c.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29821 bench.py --gpus $N --workload batched512 --steps 50 --warmup 5 2> gpurun_out/err_n8d.log | tee gpurun_out/bench_final_n8_batched512.json | cut -c1-300
tail -2 gpurun_out/err_n8d.log | cut -c1-300
