#!/bin/bash
# ncu --set full of every conv-GEMM launch of one un-graphed step (stems, Bottlenecks, transitions)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export HRF_SERIAL=1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:conv_gemm_tc -c 12 -f -o gpurun_out/r02_prof_conv_gemm python tools/profile_step.py > gpurun_out/r02_ncu_conv_gemm.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/r02_ncu_conv_gemm.log; ls -la gpurun_out/r02_prof_conv_gemm.ncu-rep
