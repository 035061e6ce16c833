#!/bin/bash
# round-2 evidence, part A (1 GPU): full GPU suite, smoke, both bench arms, other configs, sweeps
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -3 | tee $O/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee -a $O/r02_pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"
timeout 900 python bench.py > $O/r02_bench_final.json 2> $O/r02_bench_final.err; echo "bench rc=$?"; tail -2 $O/r02_bench_final.err
timeout 900 python bench.py --workload hrfuser_b_nus_r640 --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > $O/r02_bench_b_nus.json 2>/dev/null; echo "B rc=$?"
timeout 900 python bench.py --workload hrfuser_b_nus_r640 --train --batch 2 --steps 10 --warmup 3 > $O/r02_train_1gpu.json 2>/dev/null; echo "train rc=$?"
: > $O/r02_microbench_sweep.jsonl
for w in 7 14; do timeout 900 python tools/microbench.py --variant t --wins $w --iters 30 2>/dev/null | grep '"kind"' >> $O/r02_microbench_sweep.jsonl; done
timeout 900 python tools/microbench.py --variant b --wins 7 --iters 20 2>/dev/null | grep '"kind"' >> $O/r02_microbench_sweep.jsonl
wc -l $O/r02_microbench_sweep.jsonl
timeout 300 python tools/convgemm_bench.py 2>/dev/null | grep layer > $O/r02_convgemm_bench.txt; wc -l $O/r02_convgemm_bench.txt
timeout 300 python tools/neck_bench.py --json $O/r02_neck_bench.jsonl > /dev/null 2>&1; echo "neck rc=$?"
timeout 300 python tools/train_profile.py --top 30 2>/dev/null | tail -32 > $O/r02_train_profile.txt; echo "trainprof rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_final.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')}, 'e2e', d['e2e']['value'], 'raw', d['e2e_raw']['value'])
print(d['cpu_baseline']); print(d['gpu_eager_baseline']); print(d['roofline'])
print(open('gpurun_out/r02_bench_reference_arm.json').read()[:400])
for f in ('r02_bench_b_nus', 'r02_train_1gpu'):
    t = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f, t['value'], t['ms_per_step'])
PY
