#!/bin/bash
# two GPUs: multi-device / NCCL tests, bench at N=2 (c5 query-sharded, c4 with DDP)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2o_pytest_multi.log 2>&1; echo "pytest multi rc=$?" >> gpurun_out/r2o_pytest_multi.log; tail -4 gpurun_out/r2o_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err; echo "bench n2 rc=$?"
tail -c 600 gpurun_out/r2o_bench_n2.json; tail -3 gpurun_out/r2o_bench_n2.err
