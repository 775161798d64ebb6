#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
python bench.py --steps 10 --warmup 3 2> gpurun_out/err_b1.log | tee gpurun_out/bench_n1.json | cut -c1-2500
tail -3 gpurun_out/err_b1.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/err_b$N.log | tee gpurun_out/bench_n$N.json | cut -c1-2500
tail -5 gpurun_out/err_b$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-e2e 2>> gpurun_out/err_b$N.log | tee gpurun_out/bench_strong_n$N.json | cut -c1-1200
