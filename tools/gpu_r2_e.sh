#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tc_ab.py > gpurun_out/r2e_tc_ab.txt 2>&1; cat gpurun_out/r2e_tc_ab.txt
timeout 300 python tools/tc_phase_clocks.py > gpurun_out/r2e_tc_phase.txt 2>&1; tail -25 gpurun_out/r2e_tc_phase.txt
timeout 900 python -m pytest tests/test_gpu_tc_hypothesis.py tests/test_gpu_chamfer.py tests/test_gpu_layout_and_loss.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; tail -4 gpurun_out/r2e_pytest.log
