#!/bin/bash
# 2-GPU validation of the sharded operator forms (pairs / dense / ACA)
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -15) > gpurun_out/r1h_n2_pytest.log
cat gpurun_out/r1h_n2_pytest.log
