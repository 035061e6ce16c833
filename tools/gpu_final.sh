#!/bin/bash
# round-end evidence (lean): full gpu suite, smoke, both bench arms, input prologue bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
timeout 150 python tools/input_bench.py --json gpurun_out/input_bench.jsonl > /dev/null 2>&1; echo "input rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')}, 'e2e', d['e2e']['value'], d['e2e']['host_affinity_rank0'])
print(d['cpu_baseline'])
print(open('gpurun_out/bench_ref.json').read()[:300])
PY
