#!/bin/bash
# first GPU contact: op parity (all failures, not -x), then e2e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --maxfail=12 -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/ops.log
timeout 900 python -m pytest tests/test_gpu_backbone.py -m gpu -q --maxfail=6 -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/e2e.log
tail -25 gpurun_out/ops.log; tail -25 gpurun_out/e2e.log
