#!/bin/bash
# round-2 evidence on the final code: part A (suite, smoke, bench arms, sweeps) + part C (timeline, launch list)
cd "$(dirname "$0")/.."
bash tools/gpu_r2_final_a.sh
bash tools/gpu_r2_final_c.sh
