"""Policy base (crowd_sim/envs/policy/policy.py:5-49): attribute protocol shared by sim and trainable policies."""
import math


class Policy(object):
    def __init__(self):
        self.trainable = False
        self.phase = None
        self.model = None
        self.device = None
        self.last_state = None
        self.time_step = None
        self.env = None      # set when the agent is assumed to know the dynamics of the real world

    def configure(self, config):
        return

    def set_phase(self, phase):
        self.phase = phase

    def set_device(self, device):
        self.device = device

    def set_env(self, env):
        self.env = env

    def get_model(self):
        return self.model

    def predict(self, state):
        raise NotImplementedError

    @staticmethod
    def reach_destination(state):
        s = state.self_state
        return math.sqrt((s.py - s.gy) ** 2 + (s.px - s.gx) ** 2) < s.radius
