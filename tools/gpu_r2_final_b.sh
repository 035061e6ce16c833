#!/bin/bash
# round-2 evidence, part B (1 GPU): phase counters, CUPTI timeline, ncu launch list, ncu --set full captures
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
SPECS="lsa:18 mwca:18 ffn:18" CTAS="1 0" bash tools/gpu_phases.sh > $O/r02_phases.txt 2>&1; tail -3 $O/r02_phases.txt
timeout 300 python tools/trace_step.py --out $O/r02_trace.json 2>&1 | tail -1
python tools/trace_report.py $O/r02_trace.json > $O/r02_timeline.txt 2>&1; head -3 $O/r02_timeline.txt
export HRF_SERIAL=1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
   --csv --log-file $O/r02_launches.csv python tools/profile_step.py > $O/r02_ncu_list.log 2>&1; echo "list rc=$?"
for k in window_attn_v3 mixffn_v2 "conv_gemm_tc_kernel<256" "conv_gemm_tc_kernel<32"; do
  n=$(echo $k | tr -cd 'a-z0-9_')
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
     -k "regex:$k" -c 1 -f -o $O/r02_prof_$n python tools/profile_step.py > $O/r02_ncu_$n.log 2>&1; echo "$n rc=$?"
done
ls -la $O/r02_prof_*.ncu-rep
