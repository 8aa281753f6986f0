"""Import the UNMODIFIED reference (SJWang2015/AEMCARL at /root/reference) under modern numpy/torch.

Container-only tooling used by oracle/make_golden.py to produce tests/golden/*.npz.  /root/reference does not
exist on the GPU box; nothing in tests/, smoke() or bench.py imports this module at run time.

What the shim does:
  * numpy 2.x removed np.float / np.float_ / numpy.lib.arraysetops  -> aliases / stub module
  * gym, matplotlib are not installed                                -> minimal stubs
  * rvo2 (Python-RVO2) is not installed                              -> PyRVOSimulator-shaped class backed by
                                                                        our FP32 C restatement (oracle/aem_oracle.c)
  * hard-coded .cuda() calls, no GPU here                            -> identity
  * qnetwork_factory imports `utils.explorer` / `common...` relative to crowd_nav/ -> sys.path
Documented deviations when used as the oracle: model.eval() (dropout p=0.1 is otherwise live, hazard 1).
"""
import os
import sys
import types

import numpy as np

REF = os.environ.get("AEM_REFERENCE", "/root/reference")


class _PyRVOSimulator:
    """rvo2.PyRVOSimulator API subset used at orca.py:95-129; only agent 0's new velocity is solved
    (the reference reads nothing else)."""

    def __init__(self, time_step, neighbor_dist, max_neighbors, time_horizon, time_horizon_obst, radius, max_speed,
                 velocity=(0, 0)):
        self.time_step = time_step
        self.agents = []
        self.new_vel0 = (0.0, 0.0)

    def addAgent(self, pos, neighbor_dist, max_neighbors, time_horizon, time_horizon_obst, radius, max_speed,
                 velocity=(0, 0)):
        self.agents.append(dict(pos=tuple(pos), vel=tuple(velocity), pref=(0.0, 0.0), nd=neighbor_dist,
                                mn=max_neighbors, th=time_horizon, radius=radius, max_speed=max_speed))
        return len(self.agents) - 1

    def getNumAgents(self):
        return len(self.agents)

    def setAgentPosition(self, i, p):
        self.agents[i]["pos"] = tuple(p)

    def setAgentVelocity(self, i, v):
        self.agents[i]["vel"] = tuple(v)

    def setAgentPrefVelocity(self, i, v):
        self.agents[i]["pref"] = tuple(v)

    def doStep(self):
        from oracle import pyoracle
        a = self.agents[0]
        others = [[o["pos"][0], o["pos"][1], o["vel"][0], o["vel"][1], o["radius"]] for o in self.agents[1:]]
        nv = pyoracle.orca_agent(a["pos"], a["vel"], a["radius"], a["max_speed"], a["pref"],
                                 np.array(others, dtype=np.float32).reshape(-1, 5), a["nd"], a["mn"], a["th"],
                                 self.time_step)
        self.new_vel0 = (float(nv[0]), float(nv[1]))

    def getAgentVelocity(self, i):
        assert i == 0
        return self.new_vel0

    def getAgentPosition(self, i):
        return self.agents[i]["pos"]


def install():
    if getattr(install, "_done", False):
        return
    if not os.path.isdir(REF):
        raise RuntimeError("reference not present at %s (only available in the build container)" % REF)
    import torch
    np.float = float
    np.float_ = np.float64
    m = types.ModuleType("numpy.lib.arraysetops")
    m.isin = np.isin
    sys.modules["numpy.lib.arraysetops"] = m

    gym = types.ModuleType("gym")

    class Env(object):
        pass

    gym.Env = Env
    registry = {}
    envs = types.ModuleType("gym.envs")
    reg = types.ModuleType("gym.envs.registration")

    def register(id, entry_point, **kw):
        registry[id] = entry_point

    def make(id):
        import importlib
        mod, cls = registry[id].split(":")
        return getattr(importlib.import_module(mod), cls)()

    reg.register = register
    envs.registration = reg
    gym.envs = envs
    gym.make = make
    sys.modules["gym"] = gym
    sys.modules["gym.envs"] = envs
    sys.modules["gym.envs.registration"] = reg

    for name in ("matplotlib", "matplotlib.lines", "matplotlib.pyplot", "matplotlib.patches",
                 "matplotlib.animation"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].lines = sys.modules["matplotlib.lines"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    rvo2 = types.ModuleType("rvo2")
    rvo2.PyRVOSimulator = _PyRVOSimulator
    sys.modules["rvo2"] = rvo2

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self

    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (here, REF, os.path.join(REF, "crowd_nav")):
        if p in sys.path:
            sys.path.remove(p)
    # reference first so `crowd_sim` / `crowd_nav` resolve to the reference, repo root for `oracle.*`
    sys.path[:0] = [os.path.join(REF, "crowd_nav"), REF, here]
    install._done = True


def make_policy(seed=0, act_fixed=False, act_steps=1):
    """policy_factory['actenvcarl'] configured as README.md:52,56 (flag 5, self_attention), eval mode."""
    install()
    import configparser
    import contextlib
    import io
    import torch
    from crowd_nav.policy.policy_factory import policy_factory
    cfg = configparser.RawConfigParser()
    cfg.read(os.path.join(REF, "crowd_nav/configs/policy.config"))
    cfg.set("actenvcarl", "test_policy_flag", "5")
    cfg.set("actenvcarl", "multi_process", "self_attention")
    cfg.set("actenvcarl", "act_steps", act_steps)
    cfg.set("actenvcarl", "act_fixed", act_fixed)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        policy = policy_factory["actenvcarl"]()
        policy.configure(cfg)
    policy.get_model().eval()          # documented deviation (hazard 1)
    policy.set_phase("test")
    policy.set_device(torch.device("cpu"))
    return policy


def make_env(human_num=5, test_sim="circle_crossing", agent_timestep=0.4, human_timestep=0.5, reward_increment=2.0,
             position_variance=2.0, direction_variance=2.0, visible=False):
    """CrowdSim-v0 configured as test.py:115-131 with the README reward flags."""
    install()
    import configparser
    import contextlib
    import io
    import gym
    with contextlib.redirect_stdout(io.StringIO()):
        import crowd_sim  # noqa: F401  (registers CrowdSim-v0)
        from crowd_sim.envs.utils.robot import Robot
    cfg = configparser.RawConfigParser()
    cfg.read(os.path.join(REF, "crowd_nav/configs/env.config"))
    cfg.set("sim", "human_num", human_num)
    cfg.set("reward", "agent_timestep", agent_timestep)
    cfg.set("reward", "human_timestep", human_timestep)
    cfg.set("reward", "reward_increment", reward_increment)
    cfg.set("reward", "position_variance", position_variance)
    cfg.set("reward", "direction_variance", direction_variance)
    env = gym.make("CrowdSim-v0")
    env.configure(cfg)
    env.test_sim = test_sim
    env.train_val_sim = test_sim
    robot = Robot(cfg, "robot")
    robot.visible = visible
    env.set_robot(robot)
    return env, robot, cfg
