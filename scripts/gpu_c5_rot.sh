#!/bin/bash
# C5 (1000 spheres, nMax 10, FH+SH) in the rotated-axial operator form: 15 GB per harmonic, fits one GPU
set -u
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  (timeout 600 python bench.py --gpus 1 --steps 2 --warmup 2 --workload c5 --operator rot --no-cpu-baseline 2> gpurun_out/r1q_c5_rot_n1.err | tail -1) > gpurun_out/r1q_bench_c5_rot_n1.json
else
  (timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N \
     bench.py --gpus $N --steps 2 --warmup 2 --workload c5 --operator rot --no-cpu-baseline 2> gpurun_out/r1q_c5_rot_n$N.err | tail -1) > gpurun_out/r1q_bench_c5_rot_n$N.json
fi
python - <<PY
import json
d=json.load(open("gpurun_out/r1q_bench_c5_rot_n$N.json"))
print("C5 rot N=$N", d["ms_per_step"], d["e2e"]["value"], d["config"]["phases_ms_per_step"], d["roofline"]["avg_launch_ms"], d["config"]["cross_sections"], d["config"]["iters_ff"], d["config"]["iters_sh"])
PY
tail -2 gpurun_out/r1q_c5_rot_n$N.err | cut -c1-300
