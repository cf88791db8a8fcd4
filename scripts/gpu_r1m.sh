#!/bin/bash
# assembly kernel: three-case emission + 3 resident CTAs per SM; parity first, then both register bounds timed
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q > gpurun_out/r1m_parity.log 2>&1; tail -3 gpurun_out/r1m_parity.log
for b in 2 3; do
timeout 300 python bench.py --no-cpu-baseline --opt assemble_minb=$b > gpurun_out/r1m_bench_minb$b.json 2> gpurun_out/r1m_bench_minb$b.err
python - <<PY
import json
d=json.load(open("gpurun_out/r1m_bench_minb$b.json"))
print("minb=$b", d["ms_per_step"], d["config"]["phases_ms_per_step"]["assemble_ff"], d["config"]["phases_ms_per_step"]["assemble_sh"], d["assembly"]["achieved_tflops"])
PY
done
