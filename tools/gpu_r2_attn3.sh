#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "lsa or mwca or fusion or pad_mask" 2>&1 | tail -3
timeout 300 python tools/microbench.py --grids nus,stf --widths 18 --kinds lsa,mwca --mods 2 --iters 50 2>&1 | grep -v Warning | cut -c1-150
SPECS="lsa:18" CTAS="1 0" bash tools/gpu_phases.sh
