#!/usr/bin/env python
"""Probe: steady-state step time against the lane sort's "near the ground" margin (os2r_tuning.sort_margin, a pure
scheduling hint). Appends to gpurun_out/kprobe.log."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kprobe import steady, ROOT  # noqa: E402

if __name__ == '__main__':
    out = open(os.path.join(ROOT, 'gpurun_out', 'kprobe.log'), 'a')
    for mm in [float(a) for a in sys.argv[1:]] or (0.5, 1.0, 2.0, 4.0, 8.0, 20.0):
        for mode in ('fixed_hip', 'free_hip'):
            s = steady(mode=mode, pre=1200, tuning={'sort_margin': mm * 1e-3})
            print(s, flush=True)
            out.write(s + '\n')
