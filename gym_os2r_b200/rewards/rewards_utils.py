"""``tolerance``: soft indicator of "x is inside [lower, upper]" with a choice of sigmoid tails.

Host-side mirror of gym_os2r/rewards/rewards_utils.py:10-122 (itself dm_control's ``tolerance``).
The built-in reward classes are evaluated inside the CUDA step kernel; this module serves
``get_state_info`` / user-defined ``RewardBase`` subclasses and accepts numpy arrays, python
scalars, or torch tensors (batched evaluation on the GPU, no host round trip).
"""
import math

import numpy as np

_DEFAULT_VALUE_AT_MARGIN = 0.1

SIGMOIDS = ('gaussian', 'hyperbolic', 'long_tail', 'reciprocal', 'cosine', 'linear', 'quadratic',
            'tanh_squared')


def _is_torch(x):
    return type(x).__module__.split('.')[0] == 'torch'


class _Ops:
    """The handful of elementwise ops needed, dispatched on numpy vs torch."""

    def __init__(self, x):
        if _is_torch(x):
            import torch
            self.exp, self.cosh, self.cos, self.tanh, self.abs = torch.exp, torch.cosh, torch.cos, torch.tanh, torch.abs
            self.where = lambda c, a, b: torch.where(c, torch.as_tensor(a, dtype=x.dtype, device=x.device),
                                                     torch.as_tensor(b, dtype=x.dtype, device=x.device))
            self.logical_and = torch.logical_and
        else:
            self.exp, self.cosh, self.cos, self.tanh, self.abs = np.exp, np.cosh, np.cos, np.tanh, np.abs
            self.where = np.where
            self.logical_and = np.logical_and


def _sigmoids(x, value_at_1, sigmoid, ops=None):
    """1 at x == 0, ``value_at_1`` at |x| == 1, decaying with the chosen shape."""
    ops = ops or _Ops(x)
    if sigmoid in ('cosine', 'linear', 'quadratic'):
        if not 0 <= value_at_1 < 1:
            raise ValueError(f'`value_at_1` must be nonnegative and smaller than 1, got {value_at_1}.')
    elif not 0 < value_at_1 < 1:
        raise ValueError(f'`value_at_1` must be strictly between 0 and 1, got {value_at_1}.')

    if sigmoid == 'gaussian':
        scale = math.sqrt(-2 * math.log(value_at_1))
        return ops.exp(-0.5 * (x * scale) ** 2)
    if sigmoid == 'hyperbolic':
        return 1 / ops.cosh(x * math.acosh(1 / value_at_1))
    if sigmoid == 'long_tail':
        scale = math.sqrt(1 / value_at_1 - 1)
        return 1 / ((x * scale) ** 2 + 1)
    if sigmoid == 'reciprocal':
        return 1 / (ops.abs(x) * (1 / value_at_1 - 1) + 1)
    if sigmoid == 'cosine':
        sx = x * (math.acos(2 * value_at_1 - 1) / math.pi)
        return ops.where(ops.abs(sx) < 1, (1 + ops.cos(math.pi * sx)) / 2, 0.0)
    if sigmoid == 'linear':
        sx = x * (1 - value_at_1)
        return ops.where(ops.abs(sx) < 1, 1 - sx, 0.0)
    if sigmoid == 'quadratic':
        sx = x * math.sqrt(1 - value_at_1)
        return ops.where(ops.abs(sx) < 1, 1 - sx ** 2, 0.0)
    if sigmoid == 'tanh_squared':
        return 1 - ops.tanh(x * math.atanh(math.sqrt(1 - value_at_1))) ** 2
    raise ValueError(f'Unknown sigmoid type {sigmoid!r}.')


def tolerance(x, bounds=(0.0, 0.0), margin=0.0, sigmoid='gaussian',
              value_at_margin=_DEFAULT_VALUE_AT_MARGIN):
    """1 inside ``bounds``; outside, falls off with distance/margin (0 immediately if margin == 0)."""
    lower, upper = bounds
    if lower > upper:
        raise ValueError('Lower bound must be <= upper bound.')
    if margin < 0:
        raise ValueError('`margin` must be non-negative.')
    scalar = np.isscalar(x)
    xa = x if _is_torch(x) else np.asarray(x, dtype=np.float64)
    ops = _Ops(xa)
    inside = ops.logical_and(lower <= xa, xa <= upper)
    if margin == 0:
        value = ops.where(inside, 1.0, 0.0)
    else:
        d = ops.where(xa < lower, lower - xa, xa - upper) / margin
        value = ops.where(inside, 1.0, _sigmoids(d, value_at_margin, sigmoid, ops))
    return float(value) if scalar else value
