#!/bin/bash
# tensor-core attention: parity tests (bf16 paths), then bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -p no:cacheprovider --timeout 60 2>&1 | tail -40 > gpurun_out/tc_ops.log
tail -15 gpurun_out/tc_ops.log
timeout 240 python -m pytest tests/test_gpu_backbone.py -m gpu -q -x -p no:cacheprovider --timeout 120 2>&1 | tail -25 > gpurun_out/tc_e2e.log
tail -8 gpurun_out/tc_e2e.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_tc.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_tc.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('roofline', d['roofline'])
for k, v in list(d['kernels'].items())[:12]: print(k, v)
PY
# A/B: programmatic dependent launch on
HRF_PDL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('HRF_PDL=1:', round(d['value'],1), 'frames/s', d['ms_per_step'], 'ms; e2e', round(d['e2e']['value'],1))"
