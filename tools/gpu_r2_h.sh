#!/bin/bash
# ncu: launch list of the bench command, then --set full of the hot kernels from tools/prof_target.py
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/r2h_bench_plain.json 2> gpurun_out/r2h_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/r2h_ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2_launches_bench.csv
python tools/prof_target.py --emd-train > gpurun_out/r2h_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'chamfer_nn_tc_kernel|chamfer_grad_kernel|emd_auction_kernel' -c 8 -o gpurun_out/r2_prof python tools/prof_target.py --emd-train > gpurun_out/r2h_ncu_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/r2h_ncu_full.log; ls -la gpurun_out/r2_prof.ncu-rep
