#!/bin/bash
# C5 (1000 spheres, nMax 10, FH+SH: the headline configuration) on N GPUs.  Pair form: 230 GB per harmonic in total.
# N = 8, 4: both harmonics resident (57.6 / 115 GB per GPU); N = 2: one harmonic at a time (keep_matrices=0, 115 GB per
# GPU); N = 1 does not fit (230 GB > 180 GB) and fails with a clean out-of-memory error.
set -u
N=$1
mkdir -p gpurun_out
EXTRA=""
if [ "$N" = "1" ] || [ "$N" = "2" ]; then EXTRA="--opt keep_matrices=0"; fi
if [ "$N" = "1" ]; then
  (timeout 900 python bench.py --gpus 1 --steps 2 --warmup 3 --workload c5 --no-cpu-baseline $EXTRA 2> gpurun_out/r1l_c5_n1.err | tail -1) > gpurun_out/r1l_bench_c5_n1.json
else
  (timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N \
     bench.py --gpus $N --steps 2 --warmup 3 --workload c5 --no-cpu-baseline $EXTRA 2> gpurun_out/r1l_c5_n$N.err | tail -1) > gpurun_out/r1l_bench_c5_n$N.json
fi
cut -c1-1800 gpurun_out/r1l_bench_c5_n$N.json; tail -3 gpurun_out/r1l_c5_n$N.err
if [ "$N" = "8" ]; then
  (timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -5) > gpurun_out/r1l_n8_pytest.log; cat gpurun_out/r1l_n8_pytest.log
fi
