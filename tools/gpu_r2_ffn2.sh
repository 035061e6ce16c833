#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in 3 4; do echo "tile $t"; HRF_FFN_TILE=$t timeout 100 python tools/ffn_once.py --B 3 --H 50 --W 76 2>&1 | tail -2; done
mb() { timeout 300 python tools/microbench.py --grids ${G:-nus} --iters 30 --kinds mixffn --widths ${WD:-18} 2>&1 | grep '"kind"' | cut -c1-200; }
echo "== v2 12x16 288"; mb
echo "== v2 12x16 576"; HRF_FFN_TILE=3 mb
echo "== v2 12x16 384"; HRF_FFN_TILE=4 mb
bn() { timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:(v['calls'],v['avg_ms']) for k,v in d['kernels'].items() if k in ('mixffn_c18','lsa_c18')})"; }
echo "== bench v1"; HRF_FFN_V2=0 bn
echo "== bench v2 288"; bn
echo "== bench v2 576"; HRF_FFN_TILE=3 bn
