#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export HRF_SERIAL=1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:mixffn_tc_kernel -c 1 -f -o gpurun_out/prof_ffn python tools/profile_step.py > gpurun_out/ncu_ffn.log 2>&1; echo "ffn rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:window_attn_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_attn python tools/profile_step.py > gpurun_out/ncu_attn.log 2>&1; echo "attn rc=$?"
md5sum hrfuser_b200/libhrfuser_b200.so
