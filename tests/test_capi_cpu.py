"""CPU suite: the C-ABI library loads and exports every symbol include/os2r.h declares, the ctypes
mirrors have the C layout, and the product fails loudly (no CPU fallback) without a CUDA device."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from gym_os2r_b200 import _capi

from helpers import make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'os2r.h')


@pytest.fixture(scope='module')
def lib():
    if not os.path.exists(_capi.LIB_PATH):
        subprocess.check_call(['make', '-C', os.path.join(ROOT, 'gym_os2r_b200', 'csrc')])
    return _capi.load_library()


def test_every_declared_symbol_is_exported(lib):
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    declared = set(re.findall(r'\b(os2r_[a-z0-9_]+)\s*\(', text))
    assert len(declared) >= 20
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.os2r_abi_version() == _capi.ABI_VERSION


def test_ctypes_layout_matches_header():
    """Compile a tiny C program against include/os2r.h and compare sizeof / offsetof with ctypes."""
    fields = {'os2r_model': (_capi.Model, ['n_dof', 'contact_body', 'tree_R', 'mass', 'inertia', 'contact_pos', 'gravity_z', 'max_torque', 'pgs_tol']),
              'os2r_task_cfg': (_capi.TaskCfg, ['obs_dim', 'obs_kind', 'reset_laying', 'obs_low', 'done_high', 'simple_lo', 'ik_clip', 'grav_std']),
              'os2r_stats': (_capi.Stats, ['env_steps', 'sum_length']),
              'os2r_packed_layout': (_capi.PackedLayout, ['obs', 'reset_id', 'term_records', 'total_bytes', 'record_words', 'prefix_records'])}
    prints = []
    for sname, (_, fl) in fields.items():
        prints.append(f'printf("%zu\\n", sizeof({sname}));')
        prints += [f'printf("%zu\\n", offsetof({sname}, {f}));' for f in fl]
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "os2r.h"\nint main(void){' + ''.join(prints) + 'return 0;}'
    with tempfile.TemporaryDirectory() as td:
        cfile, exe = os.path.join(td, 't.c'), os.path.join(td, 't')
        open(cfile, 'w').write(src)
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), cfile, '-o', exe])
        nums = [int(x) for x in subprocess.check_output([exe]).split()]
    it = iter(nums)
    for sname, (cls, fl) in fields.items():
        assert C.sizeof(cls) == next(it), sname
        for f in fl:
            assert getattr(cls, f).offset == next(it), (sname, f)


def test_widths(lib):
    task, cm, cfg = make_config('fixed_hip')
    assert lib.os2r_state_width(C.byref(cm.struct)) == _capi.state_width(cm.struct) == 4 * 2 + 13 + 2
    assert lib.os2r_params_width(C.byref(cm.struct)) == _capi.params_width(cm.struct) == 16


def test_create_rejects_bad_arguments(lib):
    task, cm, cfg = make_config('fixed_hip')
    h = C.c_void_p()
    assert lib.os2r_create(C.byref(cm.struct), C.byref(cfg), 0, 0, 0, 1, 32, C.byref(h)) != 0
    assert b'n_envs' in lib.os2r_last_error()
    assert lib.os2r_create(C.byref(cm.struct), C.byref(cfg), 8, 0, 0, 1, 16, C.byref(h)) != 0
    assert b'precision' in lib.os2r_last_error()
    assert lib.os2r_step(None, None, None, None, None, None, None, None) != 0
    assert b'null handle' in lib.os2r_last_error()
    assert lib.os2r_step_host_packed(None, None, None, 0, None) != 0 and b'null argument' in lib.os2r_last_error()
    bad = type(cm.struct)(); C.memmove(C.byref(bad), C.byref(cm.struct), C.sizeof(bad)); bad.pgs_tol = -1.0
    assert lib.os2r_create(C.byref(bad), C.byref(cfg), 8, 0, 0, 1, 32, C.byref(h)) != 0
    assert b'pgs_tol' in lib.os2r_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_shipped_models_run_the_specialised_kernels(lib):
    """os2r_model_signature (host only): the four shipped URDFs have exactly the structure signatures the specialised
    step kernels were instantiated for (os2r_kernels.h OS2R_SHIPPED_*); a model with another structure does not, and
    runs the all-general kernels."""
    from helpers import make_config
    seen = {}
    for mode in ('simple', 'fixed', 'fixed_hip', 'free_hip'):
        task, cm, cfg = make_config(mode, reward='StraightV1' if mode == 'simple' else 'BalancingV1')
        j, c, sp = C.c_uint32(), C.c_uint32(), C.c_int32()
        assert lib.os2r_model_signature(C.byref(cm.struct), C.byref(j), C.byref(c), C.byref(sp)) == 0
        assert sp.value == 1, (mode, hex(j.value), hex(c.value))
        seen[mode] = (j.value, c.value)
    assert len(set(seen.values())) == 4
    # tilt one joint frame: no longer a shipped structure
    task, cm, cfg = make_config('fixed_hip')
    other = type(cm.struct)(); C.memmove(C.byref(other), C.byref(cm.struct), C.sizeof(other))
    other.tree_p[2][2] = 0.01
    j, c, sp = C.c_uint32(), C.c_uint32(), C.c_int32()
    assert lib.os2r_model_signature(C.byref(other), C.byref(j), C.byref(c), C.byref(sp)) == 0
    assert sp.value == 0 and j.value != seen['fixed_hip'][0]


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product path must fail loudly, never fall back to a CPU implementation."""
    task, cm, cfg = make_config('fixed_hip')
    h = C.c_void_p()
    assert lib.os2r_create(C.byref(cm.struct), C.byref(cfg), 8, 0, 0, 1, 32, C.byref(h)) != 0
    assert b'no CPU fallback' in lib.os2r_last_error()
    from gym_os2r_b200.runtimes.engine import Engine
    with pytest.raises(_capi.Os2rError):
        Engine(cm, cfg, 8)
    import gym_os2r_b200
    env = gym_os2r_b200.make('Monopod-balance-v1')
    with pytest.raises(_capi.Os2rError):
        env.reset()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under gym_os2r_b200/ may reference it."""
    pkg = os.path.join(ROOT, 'gym_os2r_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(import|from)\s+oracle\b', text, flags=re.M), f
                assert 'os2r_oracle' not in text, f
    code = 'import sys; import gym_os2r_b200; assert "oracle" not in sys.modules'
    subprocess.check_call([sys.executable, '-c', code], cwd=ROOT)
