#!/bin/bash
# round-end evidence: full gpu test-suite, smoke, bench (both arms), HRFuser-B / T-stf benches,
# win-14 sweep, launch list of the bench step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --workload hrfuser_b_nus_r640 --no-cpu-baseline --steps 30 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench B rc=$?"
timeout 300 python tools/microbench.py --wins 14 --out gpurun_out/microbench_win14.jsonl > /dev/null 2>&1; echo "micro rc=$?"
python - <<'PY'
import json
for f in ('bench.json', 'bench_b.json'):
    d = json.loads(open('gpurun_out/' + f).read().strip().splitlines()[-1])
    print(f, {k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')}, 'e2e', d['e2e']['value'])
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['cpu_baseline'])
print(open('gpurun_out/bench_ref.json').read()[:400])
PY
