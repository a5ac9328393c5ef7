"""Reset-pose inverse kinematics (reference: gym_os2r/utils/reset.py:4-40)."""
import math

_REQUIRED = ('planarizer_pitch_joint', 'upper_leg_length', 'lower_leg_length', 'central_pivot_height',
             'length_boom', 'hip_offset', 'clipping_adjust')


def leg_joint_angles(robot_def: dict):
    """Hip / knee angles that put the foot just above the ground for a boom pitch.

    ``robot_def`` holds the leg geometry in millimetres plus ``planarizer_pitch_joint`` in radians.
    Returns ``[hip, knee]``; ``[0, 0]`` when the hip is too high for the leg to reach the ground
    (triangle inequality), exactly like the reference.
    """
    extra = set(robot_def) - set(_REQUIRED)
    if extra:
        raise RuntimeError(f'unexpected keys {sorted(extra)}; allowed: {_REQUIRED}')
    pitch = robot_def['planarizer_pitch_joint']
    upper, lower = robot_def['upper_leg_length'], robot_def['lower_leg_length']
    hip_height = (robot_def['length_boom'] * math.sin(pitch) + robot_def['central_pivot_height']) / math.cos(pitch)
    reach = hip_height - robot_def['hip_offset'] - robot_def['clipping_adjust']
    if reach > upper + lower:
        return [0, 0]
    cos_hip = (upper ** 2 + reach ** 2 - lower ** 2) / (2 * upper * reach)
    hip = math.acos(max(-1.0, min(1.0, cos_hip)))
    knee = math.asin(max(-1.0, min(1.0, upper * math.sin(hip) / lower))) + hip
    return [hip, -knee]
