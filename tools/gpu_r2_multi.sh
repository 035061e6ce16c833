#!/bin/bash
# configs[2] (T-stf, 4 streams) and configs[3] (HRFuser-B SyncBN training step) on N GPUs of one box:
#   gpurun --gpus N -- bash tools/gpu_r2_multi.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
run() { if [ "$N" = 1 ]; then python "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 "$@"; fi; }
[ "$ONLY" = train ] || run bench.py --gpus $N --workload hrfuser_t_stf_r1248 --steps 30 --warmup 5 --no-cpu-baseline --no-eager-baseline 2> gpurun_out/r02_stf_${N}gpu.err | tail -1 > gpurun_out/r02_stf_${N}gpu.json
[ "$ONLY" = train ] || python -c "import json; d=json.loads(open('gpurun_out/r02_stf_${N}gpu.json').read()); print('stf', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'raw', round(d['e2e_raw']['value'],1))" || tail -5 gpurun_out/r02_stf_${N}gpu.err
run bench.py --gpus $N --workload hrfuser_b_nus_r640 --train --batch 2 --steps 10 --warmup 3 2> gpurun_out/r02_train_${N}gpu.err | tail -1 > gpurun_out/r02_train_${N}gpu.json
python -c "import json; d=json.loads(open('gpurun_out/r02_train_${N}gpu.json').read()); print('train', d['n_gpus'], round(d['value'],2), 'graph ms', round(d['ms_per_step'],1), 'eager', round(d['eager']['ms_per_step'],1), 'torch', round(d['torch_syncbn']['ms_per_step'],1))" || tail -5 gpurun_out/r02_train_${N}gpu.err
