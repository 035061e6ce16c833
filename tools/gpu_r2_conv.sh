#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_neck.py tests/test_gpu_convgemm.py -x -q -m gpu 2>&1 | tail -8
bn() { timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']//30)"; }
echo "== bench cuDNN convs"; HRF_CONVGEMM=0 bn
echo "== bench conv_gemm per stream"; HRF_STEM_LOCKSTEP=0 bn
echo "== bench conv_gemm grouped"; bn
