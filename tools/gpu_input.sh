#!/bin/bash
# input prologue evidence + regression of everything else under ABI 6:
# full gpu suite, smoke, input bench, ncu (launch list + one full capture), default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 300 python tools/input_bench.py --json gpurun_out/input_bench.jsonl 2>&1 | tail -8
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:input_prologue --csv --log-file gpurun_out/input_ncu.csv python tools/input_bench.py --once > /dev/null 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')}, 'e2e', d['e2e'])
PY
