#!/bin/bash
# BN kernels: GPU tests, HBM-throughput table, one ncu --set full capture per kernel (stem shape)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_bn_train.py -x -q -m gpu -p no:cacheprovider -s 2>&1 | grep -E 'outputs|passed|failed|Error|assert ' | head -12
timeout 120 python tools/bn_bench.py --json gpurun_out/bn_bench_fp32.jsonl > gpurun_out/bn_bench_fp32.txt 2>&1
timeout 120 python tools/bn_bench.py --dtype bf16 --json gpurun_out/bn_bench_bf16.jsonl > gpurun_out/bn_bench_bf16.txt 2>&1
tail -24 gpurun_out/bn_bench_bf16.txt
cat > /tmp/bn_once.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from hrfuser_b200 import ops
x = torch.randn(8, 64, 192, 320, device='cuda'); dy = torch.randn_like(x)
m = torch.zeros(64, device='cuda'); s = torch.ones(64, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(2):
    flush.zero_(); st = ops.bn_stats(x)
    flush.zero_(); ops.bn_normalize(x, st, s, m, 1e-5, 0.1, m.clone(), s.clone())
    flush.zero_(); sums = ops.bn_bwd_stats(x, dy, m, s)
    flush.zero_(); ops.bn_bwd_dx(x, dy, sums, st[128:], s, m, s)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bn_ --launch-skip 6 -c 6 \
  -f -o gpurun_out/r01_bn_ncu python /tmp/bn_once.py > gpurun_out/bn_ncu.log 2>&1
tail -3 gpurun_out/bn_ncu.log
