#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py -x -q -m gpu -k "mixffn or backbone" 2>&1 | tail -4
mb() { timeout 300 python tools/microbench.py --grids ${G:-nus} --iters 30 --kinds mixffn --widths ${WD:-18} 2>&1 | grep '"kind"' | cut -c1-200; }
echo "== v2 12x16 576"; mb
echo "== v2 9x16 576"; HRF_FFN_TILE=5 mb
echo "== v2 9x16 448"; HRF_FFN_TILE=6 mb
bn() { timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; }
for c in 0 4 2 1; do echo "== bench stem chunk $c"; HRF_STEM_CHUNK=$c bn; done
