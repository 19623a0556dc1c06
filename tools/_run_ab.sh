mkdir -p gpurun_out; python tools/tc_ab.py | tee gpurun_out/r2ah_tc_ab.txt
timeout 900 python -m pytest tests/test_gpu_chamfer.py tests/test_gpu_tc_hypothesis.py tests/test_reference_parity.py tests/test_gpu_guards.py -m gpu -x -q 2>&1 | tail -2
