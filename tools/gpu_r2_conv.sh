#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 0 1; do echo "== HRF_CONV_RR=$v"; HRF_CONV_RR=$v timeout 300 python tools/convgemm_bench.py 2>&1 | grep layer | cut -c1-200; done
