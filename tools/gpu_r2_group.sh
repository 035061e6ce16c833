#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "grouped or mixffn or lsa" 2>&1 | tail -2
python tools/microbench.py --grids nus --widths 18 --kinds lsa,mixffn --iters 50 2>&1 | grep kind | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['kind'], d['grid'], d['ms'])"
bash tools/gpu_r2_bench.sh HRF_LOCKSTEP=0 HRF_LOCKSTEP=1 "HRF_LOCKSTEP=0 HRF_BALANCED_GRID=0"
