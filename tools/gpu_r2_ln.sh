#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_dwconv_train.py tests/test_bn_train.py -x -q -m gpu 2>&1 | tail -8
timeout 800 python tools/train_profile.py --top 12 2>&1 | tail -13
timeout 800 python bench.py --workload hrfuser_b_nus_r640 --train --batch 2 --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('graph', d['ms_per_step'], 'eager', d['eager']['ms_per_step'], 'torch', d['torch_syncbn']['ms_per_step'], d['hrf_kernel_launches_per_step'])"
