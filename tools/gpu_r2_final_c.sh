#!/bin/bash
# round-2 evidence, part C (1 GPU): CUPTI timeline of one graph replay + ncu launch list of one step
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python tools/trace_step.py --out $O/r02_trace.json 2>&1 | tail -1
python tools/trace_report.py $O/r02_trace.json > $O/r02_timeline.txt 2>&1; head -3 $O/r02_timeline.txt
export HRF_SERIAL=1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
   --csv --log-file $O/r02_launches.csv python tools/profile_step.py > $O/r02_ncu_list.log 2>&1; echo "list rc=$?"
wc -l $O/r02_launches.csv
