#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pairs_reduce -s 20 -c 1 \
    -o gpurun_out/r1d_reduce -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1d_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_arnoldi_step -s 20 -c 1 \
    -o gpurun_out/r1d_arnoldi -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1d_b.log 2>&1
ls gpurun_out | tail -4
