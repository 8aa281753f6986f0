"""Straight-line policy (crowd_sim/envs/policy/linear.py:6-23); unused by the hot-path configs, kept for the registry."""
import math

from aemcarl_b200.crowd_sim.envs.policy.policy import Policy
from aemcarl_b200.crowd_sim.envs.utils.action import ActionXY


class Linear(Policy):
    def __init__(self):
        super().__init__()
        self.trainable = False
        self.kinematics = "holonomic"
        self.multiagent_training = True

    def predict(self, state):
        s = state.self_state
        theta = math.atan2(s.gy - s.py, s.gx - s.px)
        return ActionXY(math.cos(theta) * s.v_pref, math.sin(theta) * s.v_pref)
