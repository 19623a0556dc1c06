#!/bin/bash
# H2D bandwidth of 1.57 MB pinned copies with N GPUs copying at the same time (one process per GPU)
N=${1:-8}
for i in $(seq 0 $((N-1))); do
  CUDA_VISIBLE_DEVICES=$i python tools/h2d_probe.py > gpurun_out/h2d_$i.txt 2>&1 &
done
wait
for i in $(seq 0 $((N-1))); do echo "gpu $i: $(grep ' 1.57 MB' gpurun_out/h2d_$i.txt)  |  $(grep '100.00 MB' gpurun_out/h2d_$i.txt)"; done
nproc
