#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edges.py -m gpu -q > gpurun_out/r1k_edges.log 2>&1; grep -n "^E \|passed\|failed" gpurun_out/r1k_edges.log | cut -c1-250 | tail -30
