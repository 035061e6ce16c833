#!/bin/bash
# quick A/B bench lines (no CPU / eager baselines): value, ms/step, serial ms/step, e2e, launches per step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bn() { timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-eager-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4), 'serial', round(d['serial']['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['gpu_launches']//30)"; }
for cfg in "$@"; do
  echo "== $cfg"; env $cfg bash -c "$(declare -f bn); bn"
done
