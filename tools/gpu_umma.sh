#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_umma.py -m gpu -q -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/umma.log
