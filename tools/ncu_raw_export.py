"""Export the judged subset of an `ncu --set full` report as a small CSV (one row per metric):
duration, DRAM / L2 / L1 traffic and throughput, tensor-pipe and issue utilisation, occupancy, launch
geometry and the warp-stall breakdown.   python tools/ncu_raw_export.py report.ncu-rep out.csv"""
import csv
import subprocess
import sys

PICK = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'launch__grid_size', 'launch__block_size', 'dram__bytes_read.sum.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg', 'smsp__average_warp_latency_per_inst_issued.ratio')
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
with open(sys.argv[2], 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['launch', 'kernel', 'metric', 'unit', 'value'])
    for i, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        for k, u in zip(hdr, units):
            if k in PICK or (k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio')):
                w.writerow([i, d.get('Kernel Name', '')[:80], k, u, d[k]])
print(len(rows) - 2, 'launches ->', sys.argv[2])
