#!/bin/bash
# ncu: launch list of one serial step + full captures of the two dominant tensor-core kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export HRF_SERIAL=1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
   --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:mixffn_tc_kernel -c 2 -f -o gpurun_out/prof_ffn python tools/profile_step.py > gpurun_out/ncu_ffn.log 2>&1; echo "ffn rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:window_attn_tc_kernel -c 2 -f -o gpurun_out/prof_attn python tools/profile_step.py > gpurun_out/ncu_attn.log 2>&1; echo "attn rc=$?"
ls -la gpurun_out/*.ncu-rep
