#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_neck.py -x -q -m gpu 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "mwca" 2>&1 | tail -5
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 1500 gpurun_out/r2_bench_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_a.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print('raw', d['e2e_raw']['value'], d['e2e_raw']['h2d_bytes_per_step'])
print('eager', d['gpu_eager_baseline']); print('cpu', d['cpu_baseline']); print('roof', d['roofline'])
PY
