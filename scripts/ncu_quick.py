#!/usr/bin/env python
"""Quick look at one .ncu-rep: key metrics + opcode mix per chain-step.  usage: ncu_quick.py rep chain_steps"""
import csv, subprocess, sys
from collections import Counter
rep, steps = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
names, units, vals = rr[0], rr[1], rr[2]
for key in ('gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
            'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
            'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_pipe_fmaheavy', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'derived__smsp__inst_executed_op_local_ld', 'smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum'):
    for i, n in enumerate(names):
        if n == key:
            print(f'{key:70s} {vals[i]} {units[i]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
sr = list(csv.reader(src.splitlines()))
sh = sr[1]
isrc, iex, ist = sh.index('Source'), sh.index('Instructions Executed'), sh.index('Warp Stall Sampling (All Samples)')
ops, stalls, total, tst = Counter(), Counter(), 0, 0
for r in sr[2:]:
    if len(r) <= iex: continue
    tok = r[isrc].strip().split()
    op = (tok[1] if tok[0].startswith('@') else tok[0]).split('.')[0]
    ops[op] += int(r[iex]); stalls[op] += int(r[ist]); total += int(r[iex]); tst += int(r[ist])
print('warp instr per chain-step', round(total / steps, 2))
print(' '.join(f'{k}:{v/steps:.2f}({100*stalls[k]/max(tst,1):.0f}%)' for k, v in ops.most_common(24)))
