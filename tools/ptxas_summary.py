#!/usr/bin/env python
"""Registers / stack / spill bytes of every step_kernel instantiation from the ptxas -v log of the last build."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log = open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gym_os2r_b200', 'csrc', 'os2r_kernels.ptxas.log')).read()
for b in re.split(r"ptxas info\s+: Compiling entry function '", log)[1:]:
    name = b.split("'")[0]
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    if 'step_kernel' not in dem:
        continue
    m = re.search(r'Used (\d+) registers', b)
    sp = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', b)
    short = re.sub(r'\(.*', '', dem).replace('void os2r::', '')
    print(f'{short:55s} regs {m.group(1) if m else "?":>4s}  stack/spill-st/spill-ld {sp.groups() if sp else None}')
