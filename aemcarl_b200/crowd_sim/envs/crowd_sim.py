"""CrowdSim: the gym-style single-environment API of the reference (crowd_sim/envs/crowd_sim.py) on top of the
CUDA kernels.

configure / set_robot / reset / step / onestep_lookahead keep the reference's signatures, return types
(list[ObservableState], reward, done, info object) and error behaviour.  Scenario generation runs on the host with
numpy's global legacy stream exactly like the reference (reset seeds it per case); everything step() computes
-- the n ORCA solves, collision / goal / timeout tests and the occupancy-probability reward -- runs in the kernels
aem_orca + aem_env_lookahead with B = 1.  Agents are advanced on the host with the same Python float arithmetic
as Agent.step (agent.py:122-131).  render()/get_human_times() are visualisation helpers and out of scope.
"""
import logging

import numpy as np
import torch

from aemcarl_b200.crowd_sim.envs.utils import info as info_mod
from aemcarl_b200.crowd_sim.envs.utils.action import ActionXY
from aemcarl_b200.crowd_sim.envs.utils.agent import Human
from aemcarl_b200.crowd_sim.envs.utils.info import *  # noqa: F401,F403  (callers import the info classes from here)
from aemcarl_b200.crowd_sim.envs.utils.state import ObservableState
from aemcarl_b200 import scenarios


class CrowdSim(object):
    metadata = {"render.modes": ["human"]}

    def __init__(self):
        self.time_limit = None
        self.time_step = None
        self.robot = None
        self.humans = None
        self.global_time = None
        self.human_times = None
        self.success_reward = None
        self.collision_penalty = None
        self.discomfort_dist = None
        self.discomfort_penalty_factor = None
        self.config = None
        self.case_capacity = None
        self.case_size = None
        self.case_counter = None
        self.randomize_attributes = None
        self.train_val_sim = None
        self.test_sim = None
        self.square_width = None
        self.circle_radius = None
        self.human_num = None
        self.states = None
        self.action_values = None
        self.attention_weights = None
        self.agent_prev_vx = None
        self.agent_prev_vy = None
        self._engine = None
        self._human_actions = None

    # ------------------------------------------------------------------ configuration (crowd_sim.py:265-316)
    def configure(self, config):
        self.config = config
        self.time_limit = config.getint("env", "time_limit")
        self.time_step = config.getfloat("env", "time_step")
        self.randomize_attributes = config.getboolean("env", "randomize_attributes")
        self.success_reward = config.getfloat("reward", "success_reward")
        self.collision_penalty = config.getfloat("reward", "collision_penalty")
        self.discomfort_dist = config.getfloat("reward", "discomfort_dist")
        self.discomfort_penalty_factor = config.getfloat("reward", "discomfort_penalty_factor")
        self.stationary_penalty = config.getfloat("reward", "stationary_penalty")
        self.stationary_penalty_dist = config.getfloat("reward", "stationary_penalty_dist")
        if self.config.get("humans", "policy") == "orca":
            self.case_capacity = dict(scenarios.CASE_CAPACITY)
            self.case_size = {"train": np.iinfo(np.uint32).max - 2000, "val": config.getint("env", "val_size"),
                              "test": config.getint("env", "test_size")}
            self.train_val_sim = config.get("sim", "train_val_sim")
            self.test_sim = config.get("sim", "test_sim")
            self.square_width = config.getfloat("sim", "square_width")
            self.circle_radius = config.getfloat("sim", "circle_radius")
            self.human_num = config.getint("sim", "human_num")
        else:
            raise NotImplementedError
        self.agent_timestep = config.getfloat("reward", "agent_timestep")
        self.human_timestep = config.getfloat("reward", "human_timestep")
        self.reward_increment = config.getfloat("reward", "reward_increment")
        self.position_variance = config.getfloat("reward", "position_variance")
        self.direction_variance = config.getfloat("reward", "direction_variance")
        self.case_counter = {"train": 0, "test": 0, "val": 0}
        logging.info("human number: {}".format(self.human_num))
        logging.info("Training simulation: {}, test simulation: {}".format(self.train_val_sim, self.test_sim))
        logging.info("Square width: {}, circle width: {}".format(self.square_width, self.circle_radius))

    def set_robot(self, robot):
        self.robot = robot

    def env_params(self):
        """aem_env_params of this environment (include/aemcarl_b200.h)."""
        from aemcarl_b200.engine import make_env_params
        safety = 0.0
        if self.humans and getattr(self.humans[0].policy, "safety_space", 0):
            safety = float(self.humans[0].policy.safety_space)
        return make_env_params(time_step=self.time_step, time_limit=float(self.time_limit),
                               success_reward=self.success_reward, collision_penalty=self.collision_penalty,
                               discomfort_dist=self.discomfort_dist, agent_timestep=self.agent_timestep,
                               human_timestep=self.human_timestep, reward_increment=self.reward_increment,
                               position_variance=self.position_variance,
                               direction_variance=self.direction_variance,
                               robot_visible=bool(self.robot.visible), orca_safety_space=safety)

    # ------------------------------------------------------------------ scenario generation (crowd_sim.py:318-442)
    def _new_human(self):
        human = Human(self.config, "humans")
        return human

    def generate_random_human_position(self, human_num, rule):
        if rule not in ("square_crossing", "circle_crossing", "mixed"):
            raise ValueError("Rule doesn't exist")
        self.humans = []
        robot = scenarios._Agent(self.robot.px, self.robot.py, self.robot.gx, self.robot.gy, self.robot.radius,
                                 self.robot.v_pref)
        agents = [robot]
        proto = self._new_human()
        if rule == "mixed":        # crowd_sim.py:337-386: also redraws self.human_num
            self.human_num, drawn = scenarios._mixed_humans(np.random, robot, human_num, self.circle_radius,
                                                            self.square_width, self.discomfort_dist, proto.radius,
                                                            proto.v_pref, self.randomize_attributes)
        else:
            drawn = []
            for _ in range(human_num):
                make = scenarios._circle_human if rule == "circle_crossing" else scenarios._square_human
                extent = self.circle_radius if rule == "circle_crossing" else self.square_width
                drawn.append(make(np.random, agents, extent, self.discomfort_dist, proto.radius, proto.v_pref,
                                  self.randomize_attributes))
                agents.append(drawn[-1])
        for a in drawn:
            human = self._new_human()
            human.set(a.px, a.py, a.gx, a.gy, 0, 0, 0, radius=a.radius, v_pref=a.v_pref)
            self.humans.append(human)

    # ------------------------------------------------------------------ reset (crowd_sim.py:486-552)
    def reset(self, phase="test", test_case=None):
        if self.robot is None:
            raise AttributeError("robot has to be set!")
        assert phase in ["train", "val", "test"]
        if test_case is not None:
            self.case_counter[phase] = test_case
        self.global_time = 0
        if phase == "test":
            self.human_times = [0] * self.human_num
        else:
            self.human_times = [0] * (self.human_num if self.robot.policy.multiagent_training else 1)
        if not self.robot.policy.multiagent_training:
            self.train_val_sim = "circle_crossing"
        counter_offset = scenarios.COUNTER_OFFSET
        self.robot.set(0, -self.circle_radius, 0, self.circle_radius, 0, 0, np.pi / 2)
        self.agent_prev_vx = None
        self.agent_prev_vy = None
        if self.case_counter[phase] >= 0:
            np.random.seed(counter_offset[phase] + self.case_counter[phase])
            if phase in ["train", "val"]:
                human_num = self.human_num if self.robot.policy.multiagent_training else 1
                self.generate_random_human_position(human_num=human_num, rule=self.train_val_sim)
            else:
                self.generate_random_human_position(human_num=self.human_num, rule=self.test_sim)
            self.case_counter[phase] = (self.case_counter[phase] + 1) % self.case_size[phase]
        else:
            assert phase == "test"
            if self.case_counter[phase] == -1:
                self.human_num = 3
                self.humans = [self._new_human() for _ in range(self.human_num)]
                self.humans[0].set(0, -6, 0, 5, 0, 0, np.pi / 2)
                self.humans[1].set(-5, -5, -5, 5, 0, 0, np.pi / 2)
                self.humans[2].set(5, -5, 5, 5, 0, 0, np.pi / 2)
            else:
                raise NotImplementedError
        for agent in [self.robot] + self.humans:
            agent.time_step = self.time_step
            agent.policy.time_step = self.time_step
        self.states = list()
        if hasattr(self.robot.policy, "action_values"):
            self.action_values = list()
        if hasattr(self.robot.policy, "get_attention_weights"):
            self.attention_weights = list()
        if self.robot.sensor == "coordinates":
            return [human.get_observable_state() for human in self.humans]
        raise NotImplementedError

    # ------------------------------------------------------------------ state packing for the kernels
    def pack_state(self):
        """(robot [1,9], humans [1,n,8], global_time [1]) float64 arrays, layout of include/aemcarl_b200.h"""
        r = self.robot
        robot = np.array([[r.px, r.py, r.vx, r.vy, r.radius, r.gx, r.gy, r.v_pref, r.theta]], dtype=np.float64)
        humans = np.array([[[h.px, h.py, h.vx, h.vy, h.radius, h.gx, h.gy, h.v_pref] for h in self.humans]],
                          dtype=np.float64)
        return robot, humans, np.array([self.global_time], dtype=np.float64)

    def lookahead_all(self, engine):
        """onestep_lookahead for every action of `engine`'s action table at once (the batched form of the loop at
        multi_human_rl.py:336-344).  Returns device tensors (robot, humans_next, rewards, info, dmin, new_vel)."""
        robot, humans, gt = self.pack_state()
        dev = engine.device
        engine.set_env_params(self.env_params())
        robot_d, humans_d, gt_d = (torch.from_numpy(a).to(dev) for a in (robot, humans, gt))
        nv = engine.orca(humans_d, robot_d)
        rewards, info, dmin, hnext = engine.env_lookahead(robot_d, humans_d, nv, gt_d)
        return robot_d, hnext, rewards, info, dmin, nv

    # ------------------------------------------------------------------ step (crowd_sim.py:554-691)
    def onestep_lookahead(self, action, rebuild=True):
        return self.step(action, update=False, rebuild=rebuild)

    def step(self, action, update=True, rebuild=True):
        from aemcarl_b200.engine import get_engine
        if not isinstance(action, ActionXY):
            raise NotImplementedError("holonomic ActionXY only (the reference's step() is broken for ActionRot)")
        eng = get_engine(torch.cuda.current_device(), role="env")
        eng.set_action_space(np.array([[action.vx, action.vy]], dtype=np.float64))
        _, hnext, rewards, info, dmin, nv = self.lookahead_all(eng)
        reward = float(rewards[0, 0].item())
        code = int(info[0, 0].item())
        info_obj = info_mod.from_code(code, float(dmin[0, 0].item()))
        done = code >= 2
        if isinstance(info_obj, (info_mod.Timeout,)):
            reward = 0                                   # the reference returns the int 0 here
        nv_h = nv[0].cpu().numpy()
        human_actions = [ActionXY(float(v[0]), float(v[1])) for v in nv_h]
        if update:
            self.states.append([self.robot.get_full_state(), [human.get_full_state() for human in self.humans]])
            if hasattr(self.robot.policy, "action_values"):
                self.action_values.append(self.robot.policy.action_values)
            if hasattr(self.robot.policy, "get_attention_weights"):
                self.attention_weights.append(self.robot.policy.get_attention_weights())
            self.robot.step(action)
            for i, human_action in enumerate(human_actions):
                self.humans[i].step(human_action)
            self.global_time += self.time_step
            for i, human in enumerate(self.humans):
                if i < len(self.human_times) and self.human_times[i] == 0 and human.reached_destination():
                    self.human_times[i] = self.global_time
            ob = [human.get_observable_state() for human in self.humans]
        else:
            hn = hnext[0].cpu().numpy()
            ob = [ObservableState(float(o[0]), float(o[1]), float(o[2]), float(o[3]), float(o[4])) for o in hn]
        self.agent_prev_vx = action.vx
        self.agent_prev_vy = action.vy
        return ob, reward, done, info_obj

    def render(self, mode="human", output_file=None):
        raise NotImplementedError("visualisation is out of scope of the hot path (DESIGN.md)")
