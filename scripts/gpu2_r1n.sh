#!/bin/bash
set -u
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -k "rot" 2>&1 | tail -6) > gpurun_out/r1n_n2_pytest.log
cat gpurun_out/r1n_n2_pytest.log
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 2 --steps 3 --warmup 3 --operator rot --no-cpu-baseline 2> gpurun_out/r1n_n2_bench.err | tail -1) > gpurun_out/r1n_bench_c4_rot_n2.json
python - <<PY
import json
d=json.load(open("gpurun_out/r1n_bench_c4_rot_n2.json"))
print("rot N=2", d["ms_per_step"], d["e2e"]["value"], d["config"]["phases_ms_per_step"])
PY
