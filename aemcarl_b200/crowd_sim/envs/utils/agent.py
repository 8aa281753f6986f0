"""Agent / Robot / Human (crowd_sim/envs/utils/{agent,robot,human}.py): physical attributes, Euler kinematics
(agent.py:110-135) and the JointState hand-off to the policy."""
import math

import numpy as np

from aemcarl_b200.crowd_sim.envs.policy.policy_factory import policy_factory
from aemcarl_b200.crowd_sim.envs.utils.action import ActionRot, ActionXY
from aemcarl_b200.crowd_sim.envs.utils.state import FullState, JointState, ObservableState


class Agent(object):
    def __init__(self, config, section):
        self.visible = config.getboolean(section, "visible")
        self.v_pref = config.getfloat(section, "v_pref")
        self.radius = config.getfloat(section, "radius")
        self.policy = policy_factory[config.get(section, "policy")]()
        self.sensor = config.get(section, "sensor")
        self.kinematics = self.policy.kinematics if self.policy is not None else None
        self.px = self.py = self.gx = self.gy = self.vx = self.vy = self.theta = None
        self.time_step = None

    def print_info(self):
        import logging
        logging.info("Agent is {} and has {} kinematic constraint".format(
            "visible" if self.visible else "invisible", self.kinematics))

    def set_policy(self, policy):
        self.policy = policy
        self.kinematics = policy.kinematics

    def sample_random_attributes(self):
        self.v_pref = np.random.uniform(0.5, 1.5)      # agent.py:39-45
        self.radius = np.random.uniform(0.3, 0.5)

    def set(self, px, py, gx, gy, vx, vy, theta, radius=None, v_pref=None):
        self.px, self.py, self.gx, self.gy, self.vx, self.vy, self.theta = px, py, gx, gy, vx, vy, theta
        if radius is not None:
            self.radius = radius
        if v_pref is not None:
            self.v_pref = v_pref

    def get_observable_state(self):
        return ObservableState(self.px, self.py, self.vx, self.vy, self.radius)

    def get_next_observable_state(self, action):
        self.check_validity(action)
        px, py = self.compute_position(action, self.time_step)
        return ObservableState(px, py, action.vx, action.vy, self.radius)

    def get_full_state(self):
        return FullState(self.px, self.py, self.vx, self.vy, self.radius, self.gx, self.gy, self.v_pref, self.theta)

    def get_position(self):
        return self.px, self.py

    def set_position(self, position):
        self.px, self.py = position[0], position[1]

    def get_goal_position(self):
        return self.gx, self.gy

    def get_velocity(self):
        return self.vx, self.vy

    def set_velocity(self, velocity):
        self.vx, self.vy = velocity[0], velocity[1]

    def act(self, ob):
        raise NotImplementedError

    def check_validity(self, action):
        if self.kinematics == "holonomic":
            assert isinstance(action, ActionXY)
        else:
            assert isinstance(action, ActionRot)
            raise NotImplementedError("unicycle kinematics is broken upstream (crowd_sim.py:598); holonomic only")

    def compute_position(self, action, delta_t):
        self.check_validity(action)
        return self.px + action.vx * delta_t, self.py + action.vy * delta_t

    def step(self, action):
        self.check_validity(action)
        self.px, self.py = self.compute_position(action, self.time_step)
        self.vx, self.vy = action.vx, action.vy

    def reached_destination(self):
        return math.sqrt((self.px - self.gx) ** 2 + (self.py - self.gy) ** 2) < self.radius


class Human(Agent):
    def act(self, ob):
        return self.policy.predict(JointState(self.get_full_state(), ob))


class Robot(Agent):
    def act(self, ob):
        if self.policy is None:
            raise AttributeError("Policy attribute has to be set!")
        return self.policy.predict(JointState(self.get_full_state(), ob))
