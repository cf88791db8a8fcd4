#!/bin/bash
# 8-GPU validation: sharded parity tests (world 2/4/8), strong-scaling points on C4, and the C5 headline
# configuration (1000 spheres, nMax 10, FH+SH) on 8 GPUs.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r1_n8_gpus.txt
(timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -5) > gpurun_out/r1_n8_pytest.log
for n in 8 4 2; do
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
   bench.py --gpus $n --steps 3 --warmup 3 2>&1 | tail -1) > gpurun_out/r1_bench_c4_n$n.json
done
(timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 2>&1 | tail -1) > gpurun_out/r1_bench_c4_n1.json
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
   bench.py --gpus 8 --steps 2 --warmup 3 --workload c5 2>&1 | tail -3) > gpurun_out/r1_bench_c5_n8.json
cat gpurun_out/r1_n8_pytest.log
for f in gpurun_out/r1_bench_c4_n*.json gpurun_out/r1_bench_c5_n8.json; do echo $f; cut -c1-1500 $f; done
