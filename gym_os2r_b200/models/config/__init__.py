"""YAML-backed settings with the xpath-like get/set API of the reference
(gym_os2r/models/config/__init__.py:8-57: BaseConfig.get_config / set_config, SettingsConfig).

Differences, on purpose: anchors in the YAML file are expanded into independent containers at load
time, and ``get_config`` returns a *deep* copy (the reference returns a shallow copy, :46, so a
caller mutating a nested dict silently edits the live configuration).
"""
import copy
import os

import yaml


def _unshare(node):
    """Recursive copy WITHOUT a memo: YAML aliases (one object referenced many times) become
    independent containers (copy.deepcopy would preserve the sharing)."""
    if isinstance(node, dict):
        return {k: _unshare(v) for k, v in node.items()}
    if isinstance(node, list):
        return [_unshare(v) for v in node]
    return node


class BaseConfig:
    def __init__(self, yaml_path: str):
        here = os.path.dirname(os.path.abspath(__file__))
        with open(os.path.join(here, yaml_path)) as f:
            loaded = yaml.safe_load(f)
        # drop the anchor-holder keys ("_hip", ...) and un-share aliased containers
        self.config_dict = _unshare({k: v for k, v in loaded.items() if not k.startswith('_')})

    @staticmethod
    def _split(xpath: str):
        parts = [p for p in xpath.strip('/').split('/') if p]
        if not parts:
            raise KeyError('empty config path')
        return parts

    def set_config(self, value, xpath: str):
        """Set ``value`` at e.g. ``'resets/stand/planarizer_pitch_joint'``; missing levels are created."""
        *parents, leaf = self._split(xpath)
        node = self.config_dict
        for key in parents:
            node = node.setdefault(key, {})
        node[leaf] = value

    def get_config(self, xpath: str):
        """Return a copy of the entry at ``xpath`` (KeyError when absent, like the reference)."""
        node = self.config_dict
        for key in self._split(xpath):
            node = node[key]
        return copy.deepcopy(node)


class SettingsConfig(BaseConfig):
    """All task-mode / reset / physics settings (default file: ``default/settings.yaml``)."""

    def __init__(self, yaml_path: str = './default/settings.yaml'):
        super().__init__(yaml_path=yaml_path)
