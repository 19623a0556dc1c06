mkdir -p gpurun_out; python tools/tc_ab.py | tee gpurun_out/r2ad_tc_ab.txt
python tools/tc_phase_clocks.py | tee gpurun_out/r2ad_phase.txt | sed -n 2,9p
timeout 900 python -m pytest tests/test_gpu_chamfer.py tests/test_gpu_tc_hypothesis.py -m gpu -x -q 2>&1 | tail -2
