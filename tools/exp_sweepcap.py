#!/usr/bin/env python
"""Probe: steady-state step time of the shipped kernel against the sweep cap (os2r_model.pgs_iters) — how much of the
step is the tail of the per-block maximum sweep count. Appends to gpurun_out/kprobe.log."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kprobe import steady, ROOT  # noqa: E402

if __name__ == '__main__':
    out = open(os.path.join(ROOT, 'gpurun_out', 'kprobe.log'), 'a')
    for it in [int(a) for a in sys.argv[1:]] or (8, 6, 4, 3, 2, 1):
        s = steady(iters=it, pre=1200)
        print(s, flush=True)
        out.write(s + '\n')
