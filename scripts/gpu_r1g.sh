#!/bin/bash
# field maps: parity tests, then the whole GPU suite
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fields.py -m gpu -x -q -s > gpurun_out/r1g_fields.log 2>&1; echo "fields rc=$?" >> gpurun_out/r1g_fields.log
tail -30 gpurun_out/r1g_fields.log
