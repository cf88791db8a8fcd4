#!/bin/bash
# whole GPU suite after the ACA + field-map rows, field-map throughput, ncu captures of the new kernels
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r1h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1h_pytest.log
tail -4 gpurun_out/r1h_pytest.log
timeout 600 python scripts/field_bench.py > gpurun_out/r1h_field_bench.json 2> gpurun_out/r1h_field_bench.err; tail -c 1200 gpurun_out/r1h_field_bench.json; tail -3 gpurun_out/r1h_field_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fields -c 1 -o gpurun_out/r1h_fields -f \
    python scripts/field_bench.py --grid 128 --cpu-points 4 --repeat 1 > gpurun_out/r1h_ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_matvec_aca -s 5 -c 1 -o gpurun_out/r1h_matvec_aca -f \
    python bench.py --operator aca --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1h_ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_aca_compress -c 1 -o gpurun_out/r1h_aca_compress -f \
    python bench.py --operator aca --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1h_ncu_c.log 2>&1
ls -la gpurun_out | tail -8
