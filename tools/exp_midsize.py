"""Experiment: narrow (64-thread, up to 255 registers) against wide (224-thread, 128 registers) blocks for batches between\none wide block per SM and a full wave of them (decides step_block_threads in os2r_kernels.cu)."""
import sys, os
sys.path.insert(0, os.path.join(os.environ.get('GRAFT_REPO_ROOT', '/root/repo'), 'tools'))
from kprobe import steady
for N in (33152, 36864, 40960, 45056, 49152, 57344):
    for blk in (64, 224):
        print(steady(N=N, pre=1200, steps=150, tuning={'force_block': blk}), flush=True)
