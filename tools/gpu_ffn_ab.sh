#!/bin/bash
# A/B: depthwise conv on CUDA cores (default) vs tensor cores (HRF_FFN_DW_TC=1), CTAs/SM sweep, + ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for simt in 0 1; do for n in 1 4; do echo "== dw_tc=$simt ctas/sm=$n"; HRF_FFN_DW_TC=$simt HRF_FFN_CTAS_PER_SM=$n timeout 300 python tools/microbench.py --grids nus --iters 20 2>&1 | grep -E '"mixffn"' | grep '"C": 18' | cut -c1-160; done; done
export HRF_SERIAL=1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:mixffn_tc_kernel -c 1 -f -o gpurun_out/prof_ffn python tools/profile_step.py > gpurun_out/ncu_ffn.log 2>&1; echo "ffn rc=$?"
