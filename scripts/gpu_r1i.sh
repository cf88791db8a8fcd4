#!/bin/bash
# ACA after the persistent-scratch / chunk-count change; field-map throughput on a valid grid
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_aca.py -m gpu -x -q > gpurun_out/r1i_aca.log 2>&1; tail -3 gpurun_out/r1i_aca.log
timeout 400 python bench.py --operator aca --no-cpu-baseline > gpurun_out/r1i_bench_aca.json 2> gpurun_out/r1i_bench_aca.err; echo "bench rc=$?"
timeout 600 python scripts/field_bench.py > gpurun_out/r1i_field_bench.json 2> gpurun_out/r1i_field_bench.err; cat gpurun_out/r1i_field_bench.json; tail -3 gpurun_out/r1i_field_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fields -c 1 -o gpurun_out/r1i_fields -f \
    python scripts/field_bench.py --grid 128 --cpu-points 4 --repeat 1 > gpurun_out/r1i_ncu_a.log 2>&1
timeout 400 python bench.py --opt trace_iterations=1 --no-cpu-baseline > gpurun_out/r1i_bench_trace.json 2> gpurun_out/r1i_bench_trace.err
