#!/bin/bash
# round 2 evidence, 1 GPU: configs[4] sweep on the final kernels, configs[2] / configs[3] bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r02_microbench_sweep.jsonl
for v in t b; do for w in 7 14; do
  timeout 900 python tools/microbench.py --variant $v --wins $w --iters 20 2>/dev/null | grep '"kind"' >> gpurun_out/r02_microbench_sweep.jsonl
done; done
wc -l gpurun_out/r02_microbench_sweep.jsonl
timeout 900 python bench.py --workload hrfuser_t_stf_r1248 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_t_stf.json 2> gpurun_out/r02_bench_t_stf.err; tail -c 600 gpurun_out/r02_bench_t_stf.err; cut -c1-400 gpurun_out/r02_bench_t_stf.json
timeout 900 python bench.py --workload hrfuser_b_nus_r640 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_b_nus.json 2> gpurun_out/r02_bench_b_nus.err; tail -c 600 gpurun_out/r02_bench_b_nus.err; cut -c1-400 gpurun_out/r02_bench_b_nus.json
timeout 900 python bench.py --workload hrfuser_b_nus_r640 --train --batch 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_b_train_1gpu.json 2> gpurun_out/r02_bench_b_train.err; tail -c 600 gpurun_out/r02_bench_b_train.err; cat gpurun_out/r02_bench_b_train_1gpu.json
