#!/usr/bin/env python
"""Parity report of the SHIPPED configuration (fp32, pgs_tol 1e-6, full batch, lane-sorted blocks) against the fp64
oracle in the contact steady state -> gpurun_out/parity_report.json (copied to profiles/ per round).
    python tools/parity_report.py [fixed_hip|free_hip ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import steady_state_parity  # noqa: E402

if __name__ == '__main__':
    modes = sys.argv[1:] or ['fixed_hip', 'free_hip']
    out = [steady_state_parity(mode=m, horizons=(1, 5, 20, 100)) for m in modes]
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'parity_report.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
