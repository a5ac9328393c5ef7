#!/usr/bin/env python
"""Turn a .ncu-rep (ncu --set full) into the text summary committed under profiles/.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_step_kernel.txt ["title"]
"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__maximum_warps_per_active_cycle_pct',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__sass_inst_executed_op_local_ld.sum',
        'smsp__sass_inst_executed_op_local_st.sum', 'smsp__sass_inst_executed_op_shared_ld.sum',
        'smsp__sass_inst_executed_op_shared_st.sum', 'smsp__sass_inst_executed_op_global_ld.sum',
        'smsp__sass_inst_executed_op_global_st.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
lines = [f'# {title}', f'# source: ncu --set full --clock-control none, report {rep}', '']
for k, r in enumerate(data):
    lines.append(f'## launch {k}')
    for w in want:
        if w in col:
            lines.append(f'{w:72s} {r[col[w]]} {units[col[w]]}')
    lines.append('-- warp stall reasons (avg warps stalled per issue-active cycle) --')
    st = [(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), float(r[i]))
          for h, i in col.items() if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')]
    for name, v in sorted(st, key=lambda kv: -kv[1]):
        if v > 0.005:
            lines.append(f'   {name:28s} {v:.3f}')
    fl = {}
    for op in ('ffma', 'fmul', 'fadd', 'dfma', 'dmul', 'dadd'):
        key = f'smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed'
        if key in col:
            fl[op] = float(r[col[key]])
    if fl:
        lines.append('-- FP thread-instructions per cycle (whole GPU; peak FFMA = 18944/cycle) -- ' +
                     ', '.join(f'{k}={v:.0f}' for k, v in fl.items()))
    lines.append('')
# instruction hot spots along the SASS (250-instruction bins)
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
try:
    h = srows[1]
    d = srows[2:]
    iex, ismp = h.index('Instructions Executed'), h.index('# Samples')
    stall = {x: i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x}
    tot = sum(int(r[iex]) for r in d) or 1
    tots = sum(int(r[ismp]) for r in d) or 1
    lines.append(f'## SASS profile of launch 0: {len(d)} instructions, {tot} warp-instructions executed, {tots} stall samples')
    for c in range(0, len(d), 250):
        seg = d[c:c + 250]
        ex = sum(int(r[iex]) for r in seg)
        sm = sum(int(r[ismp]) for r in seg)
        if ex == 0 and sm == 0:
            continue
        top = sorted(((k[6:], sum(int(r[i]) for r in seg)) for k, i in stall.items()), key=lambda kv: -kv[1])[:3]
        lines.append(f'  sass[{c:5d}:{c + len(seg):5d}]  executed {ex / tot * 100:5.1f}%  samples {sm / tots * 100:5.1f}%  top stalls {top}')
except Exception as e:
    lines.append(f'(no source page: {e})')
open(out, 'w').write('\n'.join(lines) + '\n')
print('wrote', out, len(lines), 'lines')
