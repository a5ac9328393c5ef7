#!/usr/bin/env python
"""Attribute the SASS-level counters of an `ncu --set full --import-source on` report to CUDA source regions.

    python tools/ncu_by_source.py gpurun_out/r2_steady.ncu-rep [mangled-kernel-substring]

The report's source page is SASS only; `nvdisasm -g` of the cubin inside libos2r.so gives the file:line of every
instruction (built with -lineinfo). Instructions are joined by position, then grouped by named line ranges of
os2r_device.cuh / os2r_kernels.cu (the phases of one physics iteration). Prints executed warp-instructions, their share,
stall samples and the top stall reasons per region: where the time goes inside the step kernel."""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
kname = rows[0][1]
hdr = rows[1]
data = rows[2:]
m = re.search(r'step_kernel<(\w+), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(bool\)(\d), \(int\)(\d+)(?:, \(unsigned int\)(\d+), \(unsigned int\)(\d+))?>', kname)
ty = {'float': 'f', 'double': 'd', 'f2': 'NS_2f2E'}[m.group(1).split('::')[-1]]
mangled = f'step_kernelI{ty}Li{m.group(2)}ELi{m.group(3)}ELi{m.group(4)}ELb{m.group(5)}ELi{m.group(6)}E'
if m.group(7):   # structure signature (round 2b)
    mangled += f'Lj{m.group(7)}ELj{m.group(8)}E'
with tempfile.TemporaryDirectory() as td:
    subprocess.check_call(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'gym_os2r_b200', 'csrc', 'libos2r.so')], cwd=td,
                          stdout=subprocess.DEVNULL)
    dis = subprocess.run(['nvdisasm', '-g', os.path.join(td, 'os2r_kernels.sm_100a.cubin')], capture_output=True, text=True).stdout
lines = dis.splitlines()
start = next(i for i, ln in enumerate(lines) if ln.startswith('.text.') and mangled in ln)
loc, cur, func = [], ('?', 0), None
for ln in lines[start + 1:]:
    if ln.startswith('.text.') or ln.startswith('.section'):
        break
    mm = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if mm:
        cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', ln):
        loc.append(cur)
assert len(loc) >= len(data), (len(loc), len(data))
# named regions of os2r_device.cuh (line numbers looked up from marker comments so that edits do not break the tool)
src = open(os.path.join(ROOT, 'gym_os2r_b200', 'csrc', 'os2r_device.cuh')).read().splitlines()


def line_of(marker):
    return next(i + 1 for i, s in enumerate(src) if marker in s)


marks = [('helpers (sincos, rsqrt ...)', 1), ('forward pass', line_of('single forward pass: kinematics')),
         ('cholesky + qdd', line_of('Cholesky of M (and of M + dt*D')), ('joint rows set-up', line_of('constraint rows in whitened coordinates')),
         ('contact rows set-up', line_of('V Gc[NC][3][N];')), ('sweeps', line_of('projected Gauss-Seidel sweeps in whitened')),
         ('integrate', line_of('v = v* + L^-T z ; q += dt v')), ('epilogue helpers (observe, reward, reset)', line_of('task epilogue (fp64)'))]


def region(f, l):
    if f == 'os2r_kernels.cu':
        return 'kernel prologue / epilogue (os2r_kernels.cu)'
    if f != 'os2r_device.cuh':
        return f'other ({f})'
    name = marks[0][0]
    for n, l0 in marks:
        if l >= l0:
            name = n
    return name


iex, ismp, ithr = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
stall = {x[6:]: i for i, x in enumerate(hdr) if x.startswith('stall_') and 'Not Issued' not in x}
agg = {}
for r, (f, l) in zip(data, loc):
    a = agg.setdefault(region(f, l), {'n': 0, 'ex': 0, 'thr': 0, 'smp': 0, 'st': {}})
    a['n'] += 1
    a['ex'] += int(r[iex]); a['thr'] += int(r[ithr]); a['smp'] += int(r[ismp])
    for k, i in stall.items():
        a['st'][k] = a['st'].get(k, 0) + int(r[i])
tot = sum(a['ex'] for a in agg.values()) or 1
tots = sum(a['smp'] for a in agg.values()) or 1
print(f'# {kname}')
print(f'# {len(data)} SASS instructions, {tot} warp-instructions executed, {tots} stall samples')
print(f'{"region":48s} {"SASS":>6s} {"executed":>10s} {"share":>6s} {"lanes":>6s} {"samples":>8s} {"share":>6s}  top stalls')
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]['smp']):
    top = ', '.join(f'{k} {v}' for k, v in sorted(a['st'].items(), key=lambda kv: -kv[1])[:4] if v)
    print(f'{name:48s} {a["n"]:6d} {a["ex"]:10d} {a["ex"] / tot * 100:5.1f}% {a["thr"] / max(a["ex"], 1):6.1f} {a["smp"]:8d} '
          f'{a["smp"] / tots * 100:5.1f}%  {top}')
