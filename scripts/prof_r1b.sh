#!/bin/bash
# pair-operator profiling: tuning sweep (CUDA-event timings from bench.py) + full ncu captures
set -u
mkdir -p gpurun_out
for o in "pairs_kb=8" "pairs_kb=16" "pairs_kb=4" "pairs_kb=16 --opt pairs_groups=2" "pairs_kb=8 --opt pairs_groups=2" "pairs_kb=12"; do
  echo "== $o" >> gpurun_out/r1b_sweep.log
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --opt $o 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']; print(d['ms_per_step'], r['avg_launch_ms'], r['achieved'], r['frac'], d['config']['phases_ms_per_step'])" >> gpurun_out/r1b_sweep.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_matvec_pairs -s 4 -c 1 \
    -o gpurun_out/r1b_matvec_pairs -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1b_mv_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_assemble_pairs -s 1 -c 1 \
    -o gpurun_out/r1b_assemble_pairs -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1b_as_ncu.log 2>&1
cat gpurun_out/r1b_sweep.log
