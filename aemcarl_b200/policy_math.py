"""Host-side constants of the policy: the action table (crowd_nav/policy/cadrl.py:129-154).

Computed with the same numpy expressions as the reference so the table is bit-identical (tests/golden/action_rotate.npz).
"""
import itertools

import numpy as np


def build_action_space(v_pref, speed_samples=5, rotation_samples=16, kinematics="holonomic"):
    """[(0, 0)] + product(rotations, speeds), rotation-major: index = 1 + 5 * r + s.  Returns float64 [A, 2]
    (ActionXY vx, vy).  Unicycle kinematics is not supported (the reference's own step() is broken for it,
    crowd_sim.py:598 reads action.vx of an ActionRot)."""
    if kinematics != "holonomic":
        raise NotImplementedError("only holonomic kinematics (see SURVEY.md a12)")
    speeds = [(np.exp((i + 1) / speed_samples) - 1) / (np.e - 1) * v_pref for i in range(speed_samples)]
    rotations = np.linspace(0, 2 * np.pi, rotation_samples, endpoint=False)
    table = [(0.0, 0.0)]
    for rotation, speed in itertools.product(rotations, speeds):
        table.append((speed * np.cos(rotation), speed * np.sin(rotation)))
    return np.array(table, dtype=np.float64)


def speeds_rotations(v_pref, speed_samples=5, rotation_samples=16):
    speeds = [(np.exp((i + 1) / speed_samples) - 1) / (np.e - 1) * v_pref for i in range(speed_samples)]
    rotations = np.linspace(0, 2 * np.pi, rotation_samples, endpoint=False)
    return speeds, rotations
