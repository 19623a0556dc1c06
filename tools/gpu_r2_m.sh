#!/bin/bash
# final single-GPU evidence of a tree: tests, smoke, calibration, bench (both arms), ncu launch list + full capture
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log; tail -3 gpurun_out/r2m_pytest.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2m_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2m_smoke.log
timeout 900 python tools/tc_calibrate.py > gpurun_out/r2m_tc_calibrate.txt 2>&1; echo "calibrate rc=$?"; cat gpurun_out/r2m_tc_calibrate.txt
timeout 300 python tools/tc_phase_clocks.py > gpurun_out/r2m_tc_phase.txt 2>&1
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err; echo "ref arm rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2m_bench_plain.json 2> gpurun_out/r2m_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2m_launches_bench.csv python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2m_ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/prof_target.py --emd-train > gpurun_out/r2m_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'chamfer_nn_tc_kernel|chamfer_grad_kernel|emd_auction_kernel' -c 8 -f -o gpurun_out/r2m_prof python tools/prof_target.py --emd-train > gpurun_out/r2m_ncu_full.log 2>&1
echo "full rc=$?"
