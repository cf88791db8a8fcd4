#!/bin/bash
# ACA operator: parity tests, then the full GPU suite, then C4 with the compressed operator
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_aca.py -m gpu -x -q -s > gpurun_out/r1f_aca.log 2>&1; echo "aca rc=$?" >> gpurun_out/r1f_aca.log
tail -25 gpurun_out/r1f_aca.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1f_pytest.log
tail -5 gpurun_out/r1f_pytest.log
timeout 400 python bench.py --operator aca --no-cpu-baseline > gpurun_out/r1f_bench_aca.json 2> gpurun_out/r1f_bench_aca.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r1f_bench_aca.json; tail -5 gpurun_out/r1f_bench_aca.err
