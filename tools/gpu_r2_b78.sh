#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "156 or 78" 2>&1 | tail -3
for v in 1 0; do echo "== HRF_B78_SPLIT=$v"; HRF_B78_SPLIT=$v timeout 300 python tools/microbench.py --variant b --grids nus --widths 156 --kinds mixffn --iters 30 2>&1 | grep '"kind"' | cut -c1-170; done
bash tools/gpu_r2_bench.sh "HRF_B78_SPLIT=0"
timeout 600 python bench.py --workload hrfuser_b_nus_r640 --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('B', round(d['value'],1), round(d['ms_per_step'],3))"
