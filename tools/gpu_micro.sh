#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 1 2; do echo "== FFN ctas/sm=$n"; HRF_FFN_CTAS_PER_SM=$n timeout 300 python tools/microbench.py --grids nus --iters 20 2>&1 | grep -E '"mixffn"' | cut -c1-200; done
for n in 1 2 3 4; do echo "== ATTN ctas/sm=$n"; HRF_ATTN_CTAS_PER_SM=$n timeout 300 python tools/microbench.py --grids nus --iters 20 2>&1 | grep -E '"lsa"' | cut -c1-200; done
timeout 600 python tools/microbench.py --out gpurun_out/microbench.jsonl > /dev/null 2>&1; wc -l gpurun_out/microbench.jsonl
