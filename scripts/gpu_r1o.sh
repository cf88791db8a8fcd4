#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_matvec_rot -s 5 -c 1 -o gpurun_out/r1o_matvec_rot -f \
    python bench.py --operator rot --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1o_ncu.log 2>&1
ls -la gpurun_out/r1o_matvec_rot.ncu-rep
