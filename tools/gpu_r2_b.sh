#!/bin/bash
# second GPU call of round 2: 16-scanner-warp TC kernel, fused zero fill, margin fix
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
timeout 300 python tools/grad_forms.py > gpurun_out/r2b_grad_forms.txt 2>&1; tail -12 gpurun_out/r2b_grad_forms.txt
timeout 300 python tools/tc_phase_clocks.py > gpurun_out/r2b_tc_phase.txt 2>&1; tail -25 gpurun_out/r2b_tc_phase.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
