#!/bin/bash
# phase cycle counters of the tcgen05 MixFFN kernel.  Build the instrumented library first (here,
# it travels with the snapshot):
#   nvcc -std=c++17 -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 \
#     -DHRF_FFN_PROFILE -o hrfuser_b200/libhrfuser_b200_prof.so hrfuser_b200/csrc/abi.cu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in ${CTAS:-1 4}; do
  echo "== CTAs/SM = $n"
  HRF_LIB=$PWD/hrfuser_b200/libhrfuser_b200_prof.so HRF_FFN_CTAS_PER_SM=$n timeout 300 python tools/ffn_phases.py 2>&1 | tail -18
done
