#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_assemble_pairs -s 1 -c 1 \
    -o gpurun_out/r1c_assemble_pairs -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1c_as_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_arnoldi_step -s 20 -c 1 \
    -o gpurun_out/r1c_arnoldi -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1c_ar_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_matvec_pairs -s 4 -c 1 \
    -o gpurun_out/r1c_matvec_pairs -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1c_mv_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/r1c_launches_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline \
    > gpurun_out/r1c_launches_c4.log 2>&1
ls -la gpurun_out | tail -8
