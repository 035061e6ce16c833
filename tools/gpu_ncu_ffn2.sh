#!/bin/bash
# ncu full capture of the second-generation MixFFN kernel (one launch, C=18 96x160x8)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mixffn_v2 -c 1 -f \
  -o gpurun_out/prof_ffn_v2 python tools/ffn_once.py --B 8 > gpurun_out/ncu_ffn_v2.log 2>&1; echo "rc=$?"
tail -3 gpurun_out/ncu_ffn_v2.log
