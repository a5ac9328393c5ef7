#!/usr/bin/env python
"""Does the host side of a rank's D2H copy matter on this box? For a few GPUs: time a 2.5 MB pinned D2H (+ sync) with the
page-locked buffer first-touched from each NUMA node's cores (sched_setaffinity before the allocation)."""
import glob
import os
import sys
import time

import torch


def node_cpus():
    out = {}
    for p in sorted(glob.glob('/sys/devices/system/node/node*/cpulist')):
        node = int(p.split('node')[-1].split('/')[0])
        cpus = set()
        for part in open(p).read().strip().split(','):
            if not part:
                continue
            a, _, b = part.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        out[node] = cpus
    return out


nodes = node_cpus()
allowed = os.sched_getaffinity(0)
print('nodes:', {k: (min(v), max(v), len(v)) for k, v in nodes.items() if v}, 'allowed', len(allowed))
gpus = [int(x) for x in sys.argv[1:]] or [0, torch.cuda.device_count() - 1]
for g in gpus:
    torch.cuda.set_device(g)
    dev = torch.empty(2531344, dtype=torch.uint8, device=f'cuda:{g}')
    for node, cpus in nodes.items():
        tgt = cpus & allowed
        if not tgt:
            continue
        os.sched_setaffinity(0, tgt)
        host = torch.empty(2531344, dtype=torch.uint8).pin_memory()
        host.zero_()
        for _ in range(20):
            host.copy_(dev, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            host.copy_(dev, non_blocking=True)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 200
        print(f'gpu {g} buffer+thread on node {node}: {dt * 1e6:.1f} us per 2.5 MB D2H+sync = {2.531344e-3 / dt:.1f} GB/s')
        os.sched_setaffinity(0, allowed)
        del host
