#!/bin/bash
# round-1 re-entry validation: full GPU parity suite, default bench, LU timing
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r1e_gpuinfo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1e_pytest.log
tail -5 gpurun_out/r1e_pytest.log
timeout 400 python bench.py > gpurun_out/r1e_bench.json 2> gpurun_out/r1e_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r1e_bench.json
timeout 200 python scripts/lu_bench.py > gpurun_out/r1e_lu.log 2>&1; tail -8 gpurun_out/r1e_lu.log
