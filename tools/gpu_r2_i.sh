#!/bin/bash
mkdir -p gpurun_out
echo "== in-tree (vector atomics)"; timeout 300 python tools/grad_forms.py 2>&1 | grep -E "backward|step" | tee gpurun_out/r2i_grad_v2.txt
echo "== scalar atomics"; PSD_B200_LIB=$PWD/exp/libpsd_scalar_atomics.so timeout 300 python tools/grad_forms.py 2>&1 | grep -E "backward|step" | tee gpurun_out/r2i_grad_scalar.txt
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log; tail -5 gpurun_out/r2i_pytest.log
cat gpurun_out/emd_n4096_tie_order.txt
