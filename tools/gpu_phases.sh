#!/bin/bash
# Phase cycle counters of the tcgen05 kernels.  Build the instrumented library first (here; it
# travels with the snapshot and is git-ignored):
#   nvcc -std=c++17 -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 \
#     -DHRF_KERNEL_PROFILE -o hrfuser_b200/libhrfuser_b200_prof.so hrfuser_b200/csrc/abi.cu
cd "$(dirname "$0")/.."
export HRF_LIB=$PWD/hrfuser_b200/libhrfuser_b200_prof.so
for spec in ${SPECS:-ffn:18 lsa:18 lsa:36 ffn:36}; do
  set -- ${spec/:/ }
  for n in ${CTAS:-1 0}; do
    echo "== $1 C=$2 CTAs/SM=$n (0 = default)"
    HRF_FFN_CTAS_PER_SM=$n HRF_ATTN_CTAS_PER_SM=$n timeout 60 python tools/phases.py --kind $1 --C $2 2>&1 | tail -17
  done
done
