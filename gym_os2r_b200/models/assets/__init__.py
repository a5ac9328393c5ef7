"""Model files (URDF) shipped with the package; mirrors gym_os2r/models/models/__init__.py:8-62."""
import os
from typing import List

_HERE = os.path.dirname(os.path.abspath(__file__))


def get_models_path() -> str:
    return _HERE + os.sep


def get_robot_names() -> List[str]:
    return sorted(f[:-5] for f in os.listdir(_HERE) if f.endswith('.urdf'))


def get_model_file(robot_name: str) -> str:
    if robot_name not in get_robot_names():
        raise RuntimeError(f"Failed to find robot '{robot_name}'")
    return os.path.join(_HERE, robot_name + '.urdf')


def get_model_string(robot_name: str) -> str:
    with open(get_model_file(robot_name)) as f:
        return f.read()
