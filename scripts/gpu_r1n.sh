#!/bin/bash
# rotated-axial operator: parity, then C4 timing
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rot.py -m gpu -x -q > gpurun_out/r1n_rot.log 2>&1; grep -n "^E \|passed\|failed\|Error" gpurun_out/r1n_rot.log | cut -c1-260 | tail -15
timeout 300 python bench.py --no-cpu-baseline --operator rot > gpurun_out/r1n_bench_rot.json 2> gpurun_out/r1n_bench_rot.err; tail -3 gpurun_out/r1n_bench_rot.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/r1n_bench_rot.json"))
print("rot", d["ms_per_step"], d["e2e"]["value"], d["config"]["phases_ms_per_step"], d["roofline"]["avg_launch_ms"], d["config"]["cross_sections"], d["config"]["iters_ff"], d["config"]["iters_sh"])
PY
