"""ACTENVCARL policy (crowd_nav/policy/actenvcarl.py + multi_human_rl.py + cadrl.py) on the CUDA hot path.

Same public surface as the reference policy object -- configure / set_phase / set_device / set_env / set_epsilon /
get_model / predict / transform, attributes name, trainable, multiagent_training, kinematics, gamma, time_step,
last_state, action_values, action_space, speeds, rotations, use_rate_statistic -- and the same exception types.
predict() replaces the 81-iteration Python loop of multi_human_rl.py:336-372 with three kernel launches
(aem_orca, aem_env_lookahead, aem_lookahead).  predict_batch() is the vectorised addition for VecCrowdSim.
"""
import logging

import numpy as np
import torch

from aemcarl_b200 import policy_math
from aemcarl_b200.crowd_nav.common.value_network import ValueNetworkTransformerAndGRU
from aemcarl_b200.crowd_sim.envs.policy.policy import Policy
from aemcarl_b200.crowd_sim.envs.utils.action import ActionRot, ActionXY


def rotate(state, kinematics="holonomic"):
    """cadrl.py:239-274 on a [batch, 14] tensor (robot 9 + human 5) -> [batch, 13]; used by transform() for the
    replay memory.  The lookahead kernel carries its own fused copy of this arithmetic."""
    dx, dy = state[:, 5] - state[:, 0], state[:, 6] - state[:, 1]
    rot = torch.atan2(dy, dx)
    c, s = torch.cos(rot), torch.sin(rot)
    dg = torch.sqrt(dx * dx + dy * dy)
    v_pref = state[:, 7]
    vx = state[:, 2] * c + state[:, 3] * s
    vy = state[:, 3] * c - state[:, 2] * s
    radius = state[:, 4]
    theta = (state[:, 8] - rot) if kinematics == "unicycle" else torch.zeros_like(v_pref)
    vx1 = state[:, 11] * c + state[:, 12] * s
    vy1 = state[:, 12] * c - state[:, 11] * s
    px1 = (state[:, 9] - state[:, 0]) * c + (state[:, 10] - state[:, 1]) * s
    py1 = (state[:, 10] - state[:, 1]) * c - (state[:, 9] - state[:, 0]) * s
    radius1 = state[:, 13]
    ax, ay = state[:, 0] - state[:, 9], state[:, 1] - state[:, 10]
    da = torch.sqrt(ax * ax + ay * ay)
    return torch.stack([dg, v_pref, theta, radius, vx, vy, px1, py1, vx1, vy1, radius1, da, radius + radius1], dim=1)


class ACTENVCARL(Policy):
    def __init__(self):
        super().__init__()
        self.name = "ACTENVCARL"
        self.trainable = True
        self.multiagent_training = None
        self.kinematics = None
        self.epsilon = None
        self.gamma = None
        self.sampling = None
        self.speed_samples = None
        self.rotation_samples = None
        self.query_env = None
        self.action_space = None
        self.action_table = None
        self.speeds = None
        self.rotations = None
        self.action_values = None
        self.with_om = None
        self.cell_num = None
        self.cell_size = None
        self.om_channel_size = None
        self.self_state_dim = 6
        self.human_state_dim = 7
        self.joint_state_dim = self.self_state_dim + self.human_state_dim
        self.use_rate_statistic = True
        self.statistic_info = [0 for _ in range(20)]
        self.preprocess_type = "Normal"
        self._engine = None

    # ---- configuration (actenvcarl.py:21-65, cadrl.py:111-124) -------------------------------------
    def set_common_parameters(self, config):
        self.gamma = config.getfloat("rl", "gamma")
        self.kinematics = config.get("action_space", "kinematics")
        self.sampling = config.get("action_space", "sampling")
        self.speed_samples = config.getint("action_space", "speed_samples")
        self.rotation_samples = config.getint("action_space", "rotation_samples")
        self.query_env = config.getboolean("action_space", "query_env")
        self.cell_num = config.getint("om", "cell_num")
        self.cell_size = config.getfloat("om", "cell_size")
        self.om_channel_size = config.getint("om", "om_channel_size")

    def configure(self, config):
        self.set_common_parameters(config)
        ints = lambda key: [int(x) for x in config.get("actenvcarl", key).split(", ")]
        in_mlp_dims = ints("in_mlp_dims")
        sort_mlp_dims = ints("sort_mlp_dims")
        sort_mlp_attention = ints("sort_attention_dims")
        action_dims = ints("action_dims")
        self.with_om = config.getboolean("actenvcarl", "with_om")
        with_dynamic_net = config.getboolean("actenvcarl", "with_dynamic_net")
        with_global_state = config.getboolean("actenvcarl", "with_global_state")
        test_policy_flag = [int(x) for x in str(config.get("actenvcarl", "test_policy_flag")).split(", ")]
        multi_process_type = config.get("actenvcarl", "multi_process")
        act_fixed = config.get("actenvcarl", "act_fixed")
        act_steps = config.get("actenvcarl", "act_steps")
        if isinstance(act_fixed, str):
            act_fixed = act_fixed.strip().lower() in ("1", "true", "yes", "on")
        if test_policy_flag[0] != 5:
            raise NotImplementedError("only --test_policy_flag 5 (ValueNetworkTransformerAndGRU) is on the hot path")
        if self.with_om or not self.query_env:
            raise NotImplementedError("with_om / query_env=false are not used by the actenvcarl configuration")
        if self.kinematics != "holonomic":
            raise NotImplementedError("holonomic kinematics only (unicycle is broken upstream, crowd_sim.py:598)")
        self.model = ValueNetworkTransformerAndGRU(self.input_dim(), self.self_state_dim, self.joint_state_dim,
                                                   in_mlp_dims, sort_mlp_dims, sort_mlp_attention, action_dims,
                                                   with_dynamic_net, with_global_state, multi_process_type,
                                                   act_steps=int(act_steps), act_fixed=bool(act_fixed))
        self.model.eval()     # DEVIATION (SURVEY hazard 1): the reference leaves dropout p=0.1 live at inference
        self.multiagent_training = config.getboolean("actcarl", "multiagent_training")
        logging.info("Policy: %s (flag 5, self_attention) on the sm_100a lookahead kernel", self.name)

    def input_dim(self):
        return self.joint_state_dim

    def set_device(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("aemcarl_b200 runs on a CUDA device only (no CPU fallback)")
        self.model.to(self.device)

    def set_epsilon(self, epsilon):
        self.epsilon = epsilon

    def get_act_statistic_info(self):
        return self.statistic_info

    def get_attention_weights(self):
        return self.model.attention_weights

    # ---- action space (cadrl.py:129-154) -----------------------------------------------------------
    def build_action_space(self, v_pref):
        table = policy_math.build_action_space(v_pref, self.speed_samples, self.rotation_samples, self.kinematics)
        self.speeds, self.rotations = policy_math.speeds_rotations(v_pref, self.speed_samples, self.rotation_samples)
        self.action_table = table
        self.action_space = [ActionXY(float(a[0]), float(a[1])) for a in table]
        if self._engine is not None:
            self._engine.set_action_space(table)

    def propagate(self, state, action, scale=1.0):
        """cadrl.py:156-181 (holonomic)"""
        from aemcarl_b200.crowd_sim.envs.utils.state import FullState, ObservableState
        if isinstance(state, ObservableState):
            return ObservableState(state.px + action.vx * self.time_step * scale,
                                   state.py + action.vy * self.time_step * scale, action.vx, action.vy, state.radius)
        if isinstance(state, FullState):
            return FullState(state.px + action.vx * self.time_step * scale,
                             state.py + action.vy * self.time_step * scale, action.vx, action.vy, state.radius,
                             state.gx, state.gy, state.v_pref, state.theta)
        raise ValueError("Type error")

    # ---- engine plumbing ---------------------------------------------------------------------------
    def engine(self):
        from aemcarl_b200.engine import get_engine
        if self._engine is None:
            self._engine = get_engine(self.device.index or 0, role="policy")
        if self.action_table is not None and (self._engine.A != len(self.action_table)
                                              or not np.array_equal(self._engine.actions, self.action_table)):
            self._engine.set_action_space(self.action_table)
        self.model.sync_engine(self._engine)
        return self._engine

    def gamma_bar(self, v_pref):
        return pow(self.gamma, self.time_step * v_pref)

    # ---- predict (multi_human_rl.py:310-387) -------------------------------------------------------
    def predict(self, state):
        if self.phase is None or self.device is None:
            raise AttributeError("Phase, device attributes have to be set!")
        if self.phase == "train" and self.epsilon is None:
            raise AttributeError("Epsilon attribute has to be set in training phase")
        if self.reach_destination(state):
            return ActionXY(0, 0) if self.kinematics == "holonomic" else ActionRot(0, 0)
        if self.action_space is None:
            self.build_action_space(state.self_state.v_pref)
        probability = np.random.random()
        if self.phase == "train" and probability < self.epsilon:
            max_action = self.action_space[np.random.choice(len(self.action_space))]
        else:
            if self.env is None or not hasattr(self.env, "lookahead_all"):
                raise AttributeError("set_env(env) with an aemcarl_b200 CrowdSim is required (query_env=true)")
            eng = self.engine()
            robot_d, hnext, rewards, info, dmin, nv = self.env.lookahead_all(eng)
            s = state.self_state      # the robot state predict() was called with (== env.robot for Robot.act)
            robot_d = torch.tensor([[s.px, s.py, s.vx, s.vy, s.radius, s.gx, s.gy, s.v_pref, s.theta]],
                                   dtype=torch.float64, device=eng.device)
            values, net, best, steps = eng.lookahead(robot_d, hnext, rewards, self.gamma_bar(s.v_pref))
            vals = values[0].cpu().numpy()
            b = int(best[0].item())
            self.action_values = [float(v) for v in vals]
            if b < 0:
                raise ValueError("Value network is not well trained. ")
            max_action = self.action_space[b]
            if self.use_rate_statistic:
                cnt = int(steps[0, b].item())
                self.model.step_cnt = cnt
                if cnt < len(self.statistic_info):
                    self.statistic_info[cnt] += 1
        if self.phase == "train":
            self.last_state = self.transform(state)
        return max_action

    def predict_batch(self, vec_env):
        """Vectorised predict for a VecCrowdSim: returns the device tensor of selected action indices [B]."""
        eng = self.engine()
        assert vec_env.engine is eng, "VecCrowdSim must share the policy's engine (get_engine(role='policy'))"
        return vec_env.lookahead()

    def transform(self, state):
        """multi_human_rl.py:423-438: rotated joint state [n, 13] of the CURRENT state for the replay memory"""
        rows = [list(state.self_state.as_tuple()) + list(h.as_tuple()) for h in state.human_states]
        t = torch.tensor(rows, dtype=torch.float32, device=self.device)
        return rotate(t, self.kinematics)
