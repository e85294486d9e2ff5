#!/usr/bin/env python
"""Per-kernel summary of an .ncu-rep (several launches): duration, grid, registers, L2 / DRAM traffic, pipe utilisation,
issue activity and the largest stall reasons.   python scripts/ncu_kernels.py <rep> [out.json]"""
import csv, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
names, units = rr[0], rr[1]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'lts__t_sectors.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'smsp__warps_eligible.avg.per_cycle_active']
out = []
for row in rr[2:]:
    d = {'kernel': row[names.index('Kernel Name')].split('(')[0]}
    for i, n in enumerate(names):
        if n in want:
            d[n] = f'{row[i]} {units[i]}'.strip()
    stalls = {n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''): float(row[i])
              for i, n in enumerate(names) if 'issue_stalled' in n and n.endswith('_per_issue_active.ratio') and row[i]}
    d['stalls_per_issue'] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
    if 'lts__t_sectors.sum' in d:
        d['l2_traffic_bytes'] = 32 * float(d['lts__t_sectors.sum'].split()[0])
    out.append(d)
print(json.dumps(out, indent=1))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], 'w'), indent=1)
