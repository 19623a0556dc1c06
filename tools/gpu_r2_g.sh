#!/bin/bash
# compute-sanitizer, ONE tool per call: $1 = memcheck | racecheck
mkdir -p gpurun_out
timeout 300 python tools/sanitize_target.py > gpurun_out/r2g_plain_$1.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/r2g_plain_$1.log; exit 1; }
tail -2 gpurun_out/r2g_plain_$1.log
timeout 1500 compute-sanitizer --tool $1 --print-limit 50 python tools/sanitize_target.py > gpurun_out/r2g_sanitizer_$1.log 2>&1; echo "sanitizer rc=$?"
tail -25 gpurun_out/r2g_sanitizer_$1.log
