#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_loss_module.py tests/test_gpu_layout_and_loss.py -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2l_pytest.log
