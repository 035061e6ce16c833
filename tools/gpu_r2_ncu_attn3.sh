#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_attn_v3 -s 2 -c 1 -f \
  -o gpurun_out/r02_prof_attn_v3 python tools/op_once.py --kind lsa > gpurun_out/r02_ncu_attn_v3.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/r02_ncu_attn_v3.log
