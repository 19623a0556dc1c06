#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tc_ab.py > gpurun_out/r2d_tc_ab.txt 2>&1; cat gpurun_out/r2d_tc_ab.txt
