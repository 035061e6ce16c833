#!/bin/bash
# round 2: second-generation MixFFN kernel -- parity + A/B timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "mixffn" 2>&1 | tail -15
for t in 1 2; do echo "tile $t"; HRF_FFN_TILE=$t timeout 100 python tools/ffn_once.py --B 3 --H 50 --W 76 2>&1 | tail -2; done
mb() { timeout 300 python tools/microbench.py --grids ${G:-nus} --iters 30 --kinds mixffn --widths ${WD:-18} 2>&1 | grep '"kind"' | cut -c1-220; }
echo "== v1"; HRF_FFN_V2=0 mb
echo "== v2 12x16"; mb
echo "== v2 12x16 1 CTA/SM"; HRF_FFN_CTAS_PER_SM=1 mb
echo "== v2 6x16"; HRF_FFN_TILE=1 mb
echo "== v2 9x16"; HRF_FFN_TILE=2 mb
echo "== v2 stf"; G=stf mb
