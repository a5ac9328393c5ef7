#!/usr/bin/env python
"""Static instruction mix of the largest backward-branch region of one step_kernel instantiation (the physics loop plus
whatever ptxas laid out inside its address range: unlikely blocks, the rolled epilogue) — for A/B comparisons of two builds
(FP instructions against MOVs etc.), not a footprint measurement; the executed footprint comes from the ncu report
(tools/ncu_by_source.py): python tools/sass_loop_mix.py [N] [BLOCK] [MINB] [lib]"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = sys.argv[1] if len(sys.argv) > 1 else '4'
BLOCK = sys.argv[2] if len(sys.argv) > 2 else '224'
MINB = sys.argv[3] if len(sys.argv) > 3 else '2'
lib = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, 'gym_os2r_b200', 'csrc', 'libos2r.so')
names = subprocess.run(['cuobjdump', '-elf', lib], capture_output=True, text=True).stdout
fn = sorted(set(re.findall(r'_ZN4os2r11step_kernelIfLi%sELi\dELi%sELb0ELi%sE[A-Za-z0-9_]*?StatsDevE' % (N, BLOCK, MINB), names)), key=len)[0]
sass = subprocess.run(['cuobjdump', '-sass', '-fun', fn, lib], capture_output=True, text=True).stdout
ins = []
for ln in sass.splitlines():
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
bars = [i for i, (_, t) in enumerate(ins) if 'BAR.SYNC' in t]
# the loop: the last backward branch whose target precedes a barrier and spans the most instructions
best = None
for i, (addr, t) in enumerate(ins):
    m = re.search(r'BRA(?:\.U)?(?:\.\w+)*\s+(?:!?U?P\d,\s*)?`?\(?\.?L?_?x?_?\d*\)?\s*(0x[0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < addr:
            j = next(k for k, (a, _) in enumerate(ins) if a >= tgt)
            if best is None or i - j > best[1] - best[0]:
                best = (j, i)
j, i = best
body = ins[j:i + 1]
mix = Counter(re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0] for _, t in body)
print(f'{fn[:60]}...: {len(ins)} SASS instructions; physics loop {len(body)} instructions = {len(body) * 16 / 1024:.1f} KB '
      f'({sum(1 for k in bars if j <= k <= i)} barriers inside)')
fp = sum(mix[k] for k in ('FFMA', 'FMUL', 'FADD', 'FFMA2', 'FMUL2', 'FADD2'))
print('  FP32 arithmetic %d (packed %d), ' % (fp, mix['FFMA2'] + mix['FMUL2'] + mix['FADD2']) +
      ', '.join(f'{k} {v}' for k, v in mix.most_common(22)))
