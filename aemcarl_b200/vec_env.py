"""VecCrowdSim: B independent CrowdSim environments resident in HBM, stepped by the CUDA kernels.

One `decide_and_step()` is, for every environment of the batch, exactly one iteration of the reference's hot loop
(crowd_nav/utils/explorer.py:51-53):  action = robot.act(ob)  ->  MultiHumanRL.predict (81-action lookahead)
and  ob, reward, done, info = env.step(action)  (crowd_sim/envs/crowd_sim.py:557-691); finished episodes are
re-seeded from a pool of host-generated scenarios (crowd_sim.py:486-552) so the batch stays full.

State layout in HBM (float64, row-major): robot [B,9], humans [B,n,8], global_time [B] -- see include/aemcarl_b200.h.
"""
import numpy as np
import torch

from . import scenarios
from .engine import Engine, F64, I32, make_env_params


class VecCrowdSim:
    def __init__(self, engine, num_envs, human_num=5, rule="circle_crossing", phase="train", first_case=0,
                 pool_size=None, gamma=0.9, case_stride=None, scenario_kw=None, device_scenarios=False):
        assert isinstance(engine, Engine)
        self.engine = engine
        self.B, self.n = int(num_envs), int(human_num)
        self.rule, self.phase = rule, phase
        self.gamma = gamma
        self.device = engine.device
        self.pool_size = int(pool_size or num_envs)
        self.first_case = first_case
        self.case_stride = int(case_stride if case_stride is not None else num_envs)
        kw = dict(scenario_kw or {})
        self.n_active = self.pool_n_active = None
        if rule == "mixed":
            # per-environment human count (crowd_sim.py:337-386): padded rows + counts, aem_set_active_counts
            self.n = scenarios.MIXED_MAX_HUMANS
            if device_scenarios:       # aem_generate_mixed_scenarios: the same draws on the device
                self.pool_robot, self.pool_humans, self.pool_n_active = engine.generate_mixed_scenarios(
                    phase, first_case, self.pool_size, **kw)
                self.pool_robot_h = self.pool_robot[:1].cpu().numpy()
                self.pool_humans_h = None
            else:
                pr, ph, pc = scenarios.generate_mixed_pool(phase, first_case, self.pool_size, **kw)
                self.pool_robot_h, self.pool_humans_h = pr, ph
                self.pool_robot = torch.from_numpy(pr).to(self.device)
                self.pool_humans = torch.from_numpy(ph).to(self.device)
                self.pool_n_active = torch.from_numpy(pc).to(self.device)
            self.n_active = torch.zeros((self.B,), dtype=I32, device=self.device)
        elif device_scenarios:
            # aem_generate_scenarios: the same seeded MT19937 draws on the device (square rule bit-exact, circle rule
            # within an ulp of cos / sin); the host generator stays the bit-exact path of the evaluation phases
            self.pool_robot, self.pool_humans = engine.generate_scenarios(phase, first_case, self.pool_size, self.n, rule, **kw)
            self.pool_robot_h = self.pool_robot[:1].cpu().numpy()
            self.pool_humans_h = None
        else:
            pr, ph = scenarios.generate_pool(phase, first_case, self.pool_size, self.n, rule, **kw)
            self.pool_robot_h, self.pool_humans_h = pr, ph
            self.pool_robot = torch.from_numpy(pr).to(self.device)
            self.pool_humans = torch.from_numpy(ph).to(self.device)
        B, n, A = self.B, self.n, engine.A
        dev = self.device
        self.robot = torch.empty((B, 9), dtype=F64, device=dev)
        self.humans = torch.empty((B, n, 8), dtype=F64, device=dev)
        self.global_time = torch.zeros((B,), dtype=F64, device=dev)
        self.case_idx = torch.zeros((B,), dtype=I32, device=dev)
        self.new_vel = torch.empty((B, n, 2), dtype=F64, device=dev)
        self.rewards = torch.empty((B, A), dtype=F64, device=dev)
        self.info = torch.empty((B, A), dtype=I32, device=dev)
        self.dmin = torch.empty((B, A), dtype=F64, device=dev)
        self.humans_next = torch.empty((B, n, 5), dtype=F64, device=dev)
        self.values = torch.empty((B, A), dtype=F64, device=dev)
        self.net_values = torch.empty((B, A), dtype=torch.float32, device=dev)
        self.act_steps = torch.empty((B, A), dtype=torch.uint8, device=dev)
        self.best = torch.empty((B,), dtype=I32, device=dev)
        self.outcomes = torch.zeros((5,), dtype=torch.int64, device=dev)    # indexed by info code
        self.last = None
        self.reset()

    # gamma^(time_step * v_pref)  (multi_human_rl.py:366); v_pref of the robot is uniform across the batch
    def gamma_bar(self):
        v_pref = float(self.pool_robot_h[0, 7])
        return pow(self.gamma, self.engine.params.time_step * v_pref)

    def reset(self):
        """env b starts with scenario b (mod pool); its next one is b + case_stride."""
        idx = torch.arange(self.B, device=self.device) % self.pool_size
        self.robot.copy_(self.pool_robot[idx])
        self.humans.copy_(self.pool_humans[idx])
        if self.n_active is not None:
            self.n_active.copy_(self.pool_n_active[idx])
        self.global_time.zero_()
        self.case_idx.copy_((torch.arange(self.B, device=self.device) + self.case_stride).to(I32))
        self.outcomes.zero_()

    def load_cases(self, first):
        """env b plays pool scenario (first + b) mod pool from t = 0"""
        idx = (torch.arange(self.B, device=self.device) + int(first)) % self.pool_size
        self.robot.copy_(self.pool_robot[idx])
        self.humans.copy_(self.pool_humans[idx])
        if self.n_active is not None:
            self.n_active.copy_(self.pool_n_active[idx])
        self.global_time.zero_()

    def lookahead(self):
        """predict(): ORCA for the humans, per-action env lookahead, fused value lookahead, argmax."""
        e = self.engine
        e.set_active_counts(self.n_active, self.pool_n_active)          # None, None for uniform batches
        e.env_prepare(self.robot, self.humans, self.global_time,
                      out=(self.new_vel, self.rewards, self.info, self.dmin, self.humans_next))
        e.lookahead(self.robot, self.humans_next, self.rewards, self.gamma_bar(),
                    out=(self.values, self.net_values, self.best, self.act_steps))
        return self.best

    def apply(self, chosen=None, auto_reset=True):
        """env.step(action) for the chosen action index of every env; returns (reward, done, info, dmin) [B]."""
        e = self.engine
        e.set_active_counts(self.n_active, self.pool_n_active)
        chosen = self.best if chosen is None else chosen
        out = e.env_apply(self.robot, self.humans, self.new_vel, self.global_time, chosen, self.rewards, self.info,
                          self.dmin)
        if auto_reset:
            e.env_reset_done(self.robot, self.humans, self.global_time, out[1], out[2], self.pool_robot,
                             self.pool_humans, self.case_idx, self.case_stride, self.outcomes)
        self.last = out
        return out

    def decide_and_step(self, auto_reset=True):
        """lookahead() + apply() in three launches: ORCA + env lookahead, the value lookahead, argmax + env step + re-seeding"""
        e = self.engine
        e.set_active_counts(self.n_active, self.pool_n_active)
        e.env_prepare(self.robot, self.humans, self.global_time,
                      out=(self.new_vel, self.rewards, self.info, self.dmin, self.humans_next))
        e.lookahead(self.robot, self.humans_next, self.rewards, self.gamma_bar(),
                    out=(self.values, self.net_values, None, self.act_steps))
        pool = (self.pool_robot, self.pool_humans, self.case_idx, self.case_stride, self.outcomes) if auto_reset else None
        out = e.env_finish(self.values, self.robot, self.humans, self.new_vel, self.global_time, self.rewards, self.info,
                           self.dmin, best=self.best, pool=pool)
        self.last = out
        return out

    def narrow(self, n):
        """a view of the first n environments (all per-environment buffers sliced, the scenario pool shared): the kernels of a
        batch whose tail has finished run on the environments that are still playing only"""
        import copy
        v = copy.copy(self)
        v.B = int(n)
        for name in ("robot", "humans", "global_time", "case_idx", "new_vel", "rewards", "info", "dmin", "humans_next", "values",
                     "net_values", "act_steps", "best") + (("n_active",) if self.n_active is not None else ()):
            setattr(v, name, getattr(self, name)[:n])
        v.last = None
        return v

    def state_numpy(self):
        return (self.robot.cpu().numpy(), self.humans.cpu().numpy(), self.global_time.cpu().numpy())
