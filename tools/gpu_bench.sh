#!/bin/bash
# smoke + bench (both arms) + ncu launch list of one step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
   --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/smoke.log | tail -3
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('roofline', d['roofline'])
for k, v in d['kernels'].items(): print(k, v)
print('cpu', d['cpu_baseline'])
PY
