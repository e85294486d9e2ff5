#!/usr/bin/env python
"""Summarise ncu outputs into profiles/: launch list shares + key raw metrics + per-opcode instruction mix.

    python scripts/ncu_summary.py <launches.csv> <prof.ncu-rep> <tag> <chain_steps_per_launch>
"""
import csv
import json
import subprocess
import sys
from collections import Counter, defaultdict

launch_csv, rep, tag, steps_per_launch = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
out = {'tag': tag}

rows = list(csv.reader(open(launch_csv)))
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hdr]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
agg = defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg[r[ki].split('(')[0][:60]][0] += 1
        agg[r[ki].split('(')[0][:60]][1] += float(r[vi].replace(',', ''))
tot = sum(v[1] for v in agg.values())
out['launch_list'] = [{'kernel': k, 'launches': v[0], 'total_us': round(v[1] / 1e3, 1), 'share': round(v[1] / tot, 4)}
                      for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]

raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
names, units, vals = rr[0], rr[1], rr[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.sum.per_cycle_elapsed', 'sm__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg',
        'smsp__cycles_active.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__average_warp_latency_issue_stalled_short_scoreboard',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
out['metrics'] = {n: f'{vals[i]} {units[i]}' for i, n in enumerate(names)
                  if n in want or ('issue_stalled' in n and n.endswith('_per_issue_active.ratio'))
                  or n in ('smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active')}
for key in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
    pass

src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
sr = list(csv.reader(src.splitlines()))
sh = sr[1]
isrc, iex, ist = sh.index('Source'), sh.index('Instructions Executed'), sh.index('Warp Stall Sampling (All Samples)')
ops, stalls, total, tstall = Counter(), Counter(), 0, 0
for r in sr[2:]:
    if len(r) <= iex:
        continue
    txt = r[isrc].strip()
    tok = txt.split()
    op = (tok[1] if tok[0].startswith('@') else tok[0]).split('.')[0]
    ops[op] += int(r[iex])
    stalls[op] += int(r[ist])
    total += int(r[iex])
    tstall += int(r[ist])
out['warp_instructions_total'] = total
out['warp_instructions_per_chain_step'] = round(total / steps_per_launch, 2)
out['opcode_mix_per_chain_step'] = {k: {'instr': round(v / steps_per_launch, 2), 'stall_share': round(stalls[k] / max(tstall, 1), 3)}
                                    for k, v in ops.most_common(22)}
json.dump(out, open(f'profiles/{tag}.json', 'w'), indent=1)
print(json.dumps(out, indent=1)[:6000])
