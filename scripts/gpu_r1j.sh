#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fields.py -m gpu -x -q > gpurun_out/r1j_fields.log 2>&1; tail -5 gpurun_out/r1j_fields.log
timeout 600 python scripts/field_bench.py > gpurun_out/r1j_field_bench.json 2> gpurun_out/r1j_field_bench.err; cat gpurun_out/r1j_field_bench.json; tail -3 gpurun_out/r1j_field_bench.err
