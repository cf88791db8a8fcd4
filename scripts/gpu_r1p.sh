#!/bin/bash
# end-of-round validation: smoke, default bench line (with the alt operator entry), full GPU parity suite
set -u
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1p_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r1p_smoke.log | cut -c1-300
timeout 200 python bench.py > gpurun_out/r1p_bench.json 2> gpurun_out/r1p_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r1p_bench.err | cut -c1-200
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r1p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r1p_pytest.log | cut -c1-200
