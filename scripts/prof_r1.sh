#!/bin/bash
# Round-1 profiling pass (run under gpurun, ONE GPU): launch list of one bench step + full captures of the
# two dominant kernels.  Numbers printed by bench.py under ncu are NOT bench values.
set -u
mkdir -p gpurun_out
# (1) every launch of: 1 warm-up step + 1 resident step + 1 e2e step of the C4 workload
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/r1_launches_c4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline \
    > gpurun_out/r1_launches_c4.log 2>&1
# (2) the streaming matvec (dominant kernel), full set
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_matvec -s 4 -c 2 \
    -o gpurun_out/r1_matvec -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline \
    > gpurun_out/r1_matvec_ncu.log 2>&1
# (3) the block-assembly kernel, full set (one launch: it rewrites 16 GB, ncu saves/restores it per pass)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_assemble -s 1 -c 1 \
    -o gpurun_out/r1_assemble -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline \
    > gpurun_out/r1_assemble_ncu.log 2>&1
ls -la gpurun_out
