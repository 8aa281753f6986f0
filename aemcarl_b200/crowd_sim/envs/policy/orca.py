"""ORCA policy (crowd_sim/envs/policy/orca.py): same constructor attributes and predict(JointState) -> ActionXY,
solved by the sm_100a ORCA kernel (aem_orca) instead of rvo2.PyRVOSimulator.  There is no CPU path."""
import numpy as np
import torch

from aemcarl_b200.crowd_sim.envs.policy.policy import Policy
from aemcarl_b200.crowd_sim.envs.utils.action import ActionXY


class ORCA(Policy):
    def __init__(self):
        super().__init__()
        self.name = "ORCA"
        self.trainable = False
        self.multiagent_training = None
        self.kinematics = "holonomic"
        self.safety_space = 0
        self.neighbor_dist = 10
        self.max_neighbors = 10
        self.time_horizon = 5
        self.time_horizon_obst = 5
        self.radius = 0.3
        self.max_speed = 1
        self.sim = None

    def configure(self, config):
        return

    def set_phase(self, phase):
        return

    def predict(self, state):
        """orca.py:82-132: agent 0 = self (goal, v_pref), the others are observable states with zero preferred
        velocity; only agent 0's new velocity is read."""
        from aemcarl_b200.engine import get_engine, make_env_params
        eng = get_engine(torch.cuda.current_device(), role="orca")
        s = state.self_state
        m = len(state.human_states)
        humans = np.zeros((1, m + 1, 8), dtype=np.float64)
        humans[0, 0] = (s.px, s.py, s.vx, s.vy, s.radius, s.gx, s.gy, s.v_pref)
        for i, h in enumerate(state.human_states):
            humans[0, i + 1] = (h.px, h.py, h.vx, h.vy, h.radius, h.px, h.py, self.max_speed)
        robot = np.zeros((1, 9), dtype=np.float64)
        eng.set_env_params(make_env_params(time_step=self.time_step, robot_visible=False,
                                           orca_safety_space=self.safety_space))
        nv = eng.orca(torch.from_numpy(humans).to(eng.device), torch.from_numpy(robot).to(eng.device))
        v = nv[0, 0].cpu().numpy()
        self.last_state = state
        return ActionXY(float(v[0]), float(v[1]))
