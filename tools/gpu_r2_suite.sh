#!/bin/bash
# full GPU suite + one bench line + launch list (round-2 checkpoint)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -c 1500 gpurun_out/r02_bench.json
