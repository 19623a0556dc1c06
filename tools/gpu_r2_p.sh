#!/bin/bash
# four GPUs: bench at N=4
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2p_bench_n4.json 2> gpurun_out/r2p_bench_n4.err; echo "bench n4 rc=$?"
tail -2 gpurun_out/r2p_bench_n4.err
