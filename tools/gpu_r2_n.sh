#!/bin/bash
# eight GPUs: bench at N=8 (weak scaling of the headline, c5 query-sharded, c4 with DDP)
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2n_bench_n8.json 2> gpurun_out/r2n_bench_n8.err; echo "bench n8 rc=$?"
tail -c 600 gpurun_out/r2n_bench_n8.json; tail -3 gpurun_out/r2n_bench_n8.err
