"""Monopod task with RAW (un-normalised) observations — mirror of
gym_os2r/tasks/monopod_no_norm.py:15-348. Identical to ``tasks.monopod.MonopodTask`` except that
the observation stops after the periodic wrap (:241-246), spaces keep the raw limits (:167-182)
and the reward class is built with ``normalized=False`` (:174)."""
from .monopod import MonopodTask as _NormalisedTask


class MonopodTask(_NormalisedTask):
    normalized = False
