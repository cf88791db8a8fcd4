#!/bin/bash
# 2-GPU validation: sharded parity test + strong-scaling bench points N=1,2 on the C4 workload
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r1_n2_gpus.txt
(timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -15) > gpurun_out/r1_n2_pytest.log
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -3) > gpurun_out/r1_n2_bench.log
cat gpurun_out/r1_n2_pytest.log gpurun_out/r1_n2_bench.log
