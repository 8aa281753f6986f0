"""Engine: one aem_handle on one CUDA device; torch owns device memory and streams (plumbing only), every
computation is a kernel of libaemcarl_b200.so launched on torch's current stream."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import EnvParams, check

F64, F32, I32, U8 = torch.float64, torch.float32, torch.int32, torch.uint8


def make_env_params(time_step=0.2, time_limit=20.0, success_reward=1.0, collision_penalty=-0.25,
                    discomfort_dist=0.2, agent_timestep=0.4, human_timestep=0.5, reward_increment=2.0,
                    position_variance=2.0, direction_variance=2.0, robot_visible=False, orca_safety_space=0.0):
    """env.config defaults with the README reward flags (README.md:56, test.py:36-40)."""
    return EnvParams(time_step, time_limit, success_reward, collision_penalty, discomfort_dist, agent_timestep,
                     human_timestep, reward_increment, position_variance, direction_variance,
                     int(bool(robot_visible)), 0, orca_safety_space)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class Engine:
    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise _lib.AemError("aemcarl_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)          # make sure the primary context exists
        h = ctypes.c_void_p()
        check(self.lib.aem_create(self.device.index, ctypes.byref(h)))
        self.h = h
        self.A = 0
        self.params = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.aem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration -------------------------------------------------------------------------
    def load_weights(self, flat):
        flat = np.ascontiguousarray(flat, dtype=np.float32)
        check(self.lib.aem_load_weights(self.h, flat.ctypes.data_as(ctypes.c_void_p), flat.size))

    def set_action_space(self, actions):
        actions = np.ascontiguousarray(actions, dtype=np.float64).reshape(-1, 2)
        check(self.lib.aem_set_action_space(self.h, actions.ctypes.data_as(ctypes.c_void_p), actions.shape[0]))
        self.A = actions.shape[0]
        self.actions = actions

    def set_env_params(self, params):
        check(self.lib.aem_set_env_params(self.h, ctypes.byref(params)))
        self.params = params

    def set_act_mode(self, act_fixed, act_steps):
        check(self.lib.aem_set_act_mode(self.h, int(bool(act_fixed)), int(act_steps)))

    def set_precision(self, precision):
        """'fp32' (FFMA kernel, default: the parity anchor), 'tf32x3' (tcgen05, one tile / SM, FP32-range containers),
        'fp16x3' (tcgen05, two tiles / SM: bench.py's default, FP32-grade) or 'fp16x1' (the same kernel with ONE fp16 pass per
        product: opt-in, values within ~1e-2 relative -- its own tolerance, tests/test_gpu_parity.py)"""
        mode = {"fp32": 0, "tf32x3": 1, "fp16x3": 2, "fp16x1": 3, 0: 0, 1: 1, 2: 2, 3: 3}[precision]
        check(self.lib.aem_set_precision(self.h, mode))
        self.precision = mode

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, t, dtype, shape=None):
        assert t.is_cuda and t.dtype == dtype and t.is_contiguous(), (t.device, t.dtype)
        if shape is not None:
            assert tuple(t.shape) == tuple(shape), (tuple(t.shape), shape)
        return t

    # ---- device entry points -------------------------------------------------------------------
    def orca(self, humans, robot, out=None):
        B, n, _ = humans.shape
        self._chk(humans, F64, (B, n, 8)); self._chk(robot, F64, (B, 9))
        if out is None:
            out = torch.empty((B, n, 2), dtype=F64, device=self.device)
        check(self.lib.aem_orca(self.h, B, n, _ptr(humans), _ptr(robot), _ptr(out), self._stream()))
        return out

    def env_lookahead(self, robot, humans, new_vel, global_time, out=None):
        B, n, _ = humans.shape
        A = self.A
        self._chk(humans, F64, (B, n, 8)); self._chk(robot, F64, (B, 9))
        self._chk(new_vel, F64, (B, n, 2)); self._chk(global_time, F64, (B,))
        if out is None:
            out = (torch.empty((B, A), dtype=F64, device=self.device), torch.empty((B, A), dtype=I32, device=self.device),
                   torch.empty((B, A), dtype=F64, device=self.device),
                   torch.empty((B, n, 5), dtype=F64, device=self.device))
        rewards, info, dmin, hnext = out
        check(self.lib.aem_env_lookahead(self.h, B, n, _ptr(robot), _ptr(humans), _ptr(new_vel), _ptr(global_time),
                                         _ptr(rewards), _ptr(info), _ptr(dmin), _ptr(hnext), self._stream()))
        return rewards, info, dmin, hnext

    def env_prepare(self, robot, humans, global_time, out=None):
        """aem_orca + aem_env_lookahead in one launch -> (new_vel, rewards, info, dmin, humans_next)"""
        B, n, _ = humans.shape
        A = self.A
        self._chk(humans, F64, (B, n, 8)); self._chk(robot, F64, (B, 9)); self._chk(global_time, F64, (B,))
        if out is None:
            out = (torch.empty((B, n, 2), dtype=F64, device=self.device),
                   torch.empty((B, A), dtype=F64, device=self.device), torch.empty((B, A), dtype=I32, device=self.device),
                   torch.empty((B, A), dtype=F64, device=self.device),
                   torch.empty((B, n, 5), dtype=F64, device=self.device))
        new_vel, rewards, info, dmin, hnext = out
        check(self.lib.aem_env_prepare(self.h, B, n, _ptr(robot), _ptr(humans), _ptr(global_time), _ptr(new_vel),
                                       _ptr(rewards), _ptr(info), _ptr(dmin), _ptr(hnext), self._stream()))
        return new_vel, rewards, info, dmin, hnext

    def env_finish(self, values, robot, humans, new_vel, global_time, rewards, info, dmin, best=None, pool=None):
        """argmax + aem_env_apply (+ aem_env_reset_done when `pool` = (pool_robot, pool_humans, case_idx, case_stride,
        outcome_counts)) in one launch -> (reward, done, info, dmin) [B]; `best` [B] i32 receives the chosen actions"""
        B, n, _ = humans.shape
        self._chk(values, F64, (B, self.A))
        o_reward = torch.empty((B,), dtype=F64, device=self.device)
        o_done = torch.empty((B,), dtype=I32, device=self.device)
        o_info = torch.empty((B,), dtype=I32, device=self.device)
        o_dmin = torch.empty((B,), dtype=F64, device=self.device)
        if pool is not None:
            pool_robot, pool_humans, case_idx, case_stride, outcome_counts = pool
            P = pool_robot.shape[0]
            self._chk(case_idx, I32, (B,)); self._chk(pool_robot, F64, (P, 9)); self._chk(pool_humans, F64, (P, n, 8))
        else:
            pool_robot = pool_humans = case_idx = outcome_counts = None
            P = case_stride = 0
        check(self.lib.aem_env_finish(self.h, B, n, _ptr(values), _ptr(robot), _ptr(humans), _ptr(new_vel), _ptr(global_time),
                                      _ptr(rewards), _ptr(info), _ptr(dmin), _ptr(best), _ptr(o_reward), _ptr(o_done),
                                      _ptr(o_info), _ptr(o_dmin), _ptr(pool_robot), _ptr(pool_humans), int(P),
                                      _ptr(case_idx), int(case_stride), _ptr(outcome_counts), self._stream()))
        return o_reward, o_done, o_info, o_dmin

    def lookahead(self, robot, humans_next, rewards, gamma_bar, out=None, want_net=True):
        B, n, _ = humans_next.shape
        A = self.A
        self._chk(robot, F64, (B, 9)); self._chk(humans_next, F64, (B, n, 5)); self._chk(rewards, F64, (B, A))
        if out is None:
            out = (torch.empty((B, A), dtype=F64, device=self.device),
                   torch.empty((B, A), dtype=F32, device=self.device) if want_net else None,
                   torch.empty((B,), dtype=I32, device=self.device),
                   torch.empty((B, A), dtype=U8, device=self.device) if want_net else None)
        values, net, best, steps = out
        check(self.lib.aem_lookahead(self.h, B, n, _ptr(robot), _ptr(humans_next), _ptr(rewards), float(gamma_bar),
                                     _ptr(values), _ptr(net), _ptr(best), _ptr(steps), self._stream()))
        return values, net, best, steps

    def env_apply(self, robot, humans, new_vel, global_time, chosen, rewards=None, info=None, dmin=None):
        B, n, _ = humans.shape
        self._chk(chosen, I32, (B,))
        o_reward = torch.empty((B,), dtype=F64, device=self.device)
        o_done = torch.empty((B,), dtype=I32, device=self.device)
        o_info = torch.empty((B,), dtype=I32, device=self.device)
        o_dmin = torch.empty((B,), dtype=F64, device=self.device)
        check(self.lib.aem_env_apply(self.h, B, n, _ptr(robot), _ptr(humans), _ptr(new_vel), _ptr(global_time),
                                     _ptr(chosen), _ptr(rewards), _ptr(info), _ptr(dmin), _ptr(o_reward),
                                     _ptr(o_done), _ptr(o_info), _ptr(o_dmin), self._stream()))
        return o_reward, o_done, o_info, o_dmin

    def env_reset_done(self, robot, humans, global_time, done, info, pool_robot, pool_humans, case_idx, case_stride,
                       outcome_counts=None):
        B, n, _ = humans.shape
        P = pool_robot.shape[0]
        self._chk(done, I32, (B,)); self._chk(case_idx, I32, (B,))
        self._chk(pool_robot, F64, (P, 9)); self._chk(pool_humans, F64, (P, n, 8))
        check(self.lib.aem_env_reset_done(self.h, B, n, _ptr(robot), _ptr(humans), _ptr(global_time), _ptr(done),
                                          _ptr(info), _ptr(pool_robot), _ptr(pool_humans), P, _ptr(case_idx),
                                          int(case_stride), _ptr(outcome_counts), self._stream()))

    def value_forward(self, states):
        S, n, d = states.shape
        self._chk(states, F32, (S, n, 13))
        values = torch.empty((S,), dtype=F32, device=self.device)
        steps = torch.empty((S,), dtype=U8, device=self.device)
        check(self.lib.aem_value_forward(self.h, S, n, _ptr(states), _ptr(values), _ptr(steps), self._stream()))
        return values, steps

    def gru_act(self, x, T):
        rows, d = x.shape
        assert d == 13 and rows % T == 0
        self._chk(x, F32)
        S = rows // T
        out = torch.empty((rows, 50), dtype=F32, device=self.device)
        steps = torch.empty((S,), dtype=U8, device=self.device)
        check(self.lib.aem_gru_act(self.h, S, T, _ptr(x), _ptr(out), _ptr(steps), self._stream()))
        return out, steps

    def env_step_xy(self, robot, humans, new_vel, global_time, actions, max_speed):
        """CrowdSim.step with one arbitrary ActionXY per env (actions [B,2] f64); updates the state in place and returns
        (reward f64, done i32, info i32, dmin f64), each [B]"""
        B, n, _ = humans.shape
        for t in (robot, humans, new_vel, global_time, actions):
            self._chk(t, F64)
        dev = self.device
        scratch = torch.empty((3 * B,), dtype=F64, device=dev)
        reward = torch.empty((B,), dtype=F64, device=dev)
        done = torch.empty((B,), dtype=I32, device=dev)
        info = torch.empty((B,), dtype=I32, device=dev)
        dmin = torch.empty((B,), dtype=F64, device=dev)
        check(self.lib.aem_env_step_xy(self.h, B, n, _ptr(robot), _ptr(humans), _ptr(new_vel), _ptr(global_time),
                                       _ptr(actions), float(max_speed), _ptr(scratch), _ptr(reward), _ptr(done),
                                       _ptr(info), _ptr(dmin), self._stream()))
        return reward, done, info, dmin

    def transform(self, robot, humans, out=None):
        """MultiHumanRL.transform of every env: robot [B,9] f64, humans [B,n,8] f64 -> rotated states [B,n,13] f32"""
        B, n, _ = humans.shape
        self._chk(robot, F64); self._chk(humans, F64)
        out = torch.empty((B, n, 13), dtype=F32, device=self.device) if out is None else out
        check(self.lib.aem_transform(self.h, B, n, _ptr(robot), _ptr(humans), _ptr(out), self._stream()))
        return out

    def generate_scenarios(self, phase, first_case, count, human_num, rule, circle_radius=4.0, square_width=10.0,
                           discomfort_dist=0.2, robot_radius=0.3, robot_v_pref=1.0, human_radius=0.3, human_v_pref=1.0,
                           randomize_attributes=False):
        """CrowdSim.reset's scenario draw for cases first_case .. first_case + count - 1 of `phase` on the device; the same
        arguments and layout as scenarios.generate_pool -> (robot [count, 9], humans [count, n, 8]) f64 device tensors"""
        from . import scenarios
        rules = {"circle_crossing": 0, "square_crossing": 1}
        if rule not in rules:
            raise NotImplementedError("device scenario generation covers circle_crossing and square_crossing")
        p = _lib.ScenarioParams(circle_radius, square_width, discomfort_dist, robot_radius, robot_v_pref, human_radius,
                                human_v_pref, int(bool(randomize_attributes)), 0)
        robot = torch.empty((count, 9), dtype=F64, device=self.device)
        humans = torch.empty((count, human_num, 8), dtype=F64, device=self.device)
        seed = scenarios.case_seed(phase, int(first_case))
        check(self.lib.aem_generate_scenarios(self.h, rules[rule], ctypes.c_uint32(seed & 0xffffffff), int(count), int(human_num),
                                              ctypes.byref(p), _ptr(robot), _ptr(humans), self._stream()))
        return robot, humans

    def generate_mixed_scenarios(self, phase, first_case, count, circle_radius=4.0, square_width=10.0, discomfort_dist=0.2,
                                 robot_radius=0.3, robot_v_pref=1.0, human_radius=0.3, human_v_pref=1.0,
                                 randomize_attributes=False):
        """the 'mixed' rule on the device; the layout of scenarios.generate_mixed_pool -> (robot [count, 9], humans
        [count, 5, 8] padded, n_active [count] int32) device tensors"""
        from . import scenarios
        p = _lib.ScenarioParams(circle_radius, square_width, discomfort_dist, robot_radius, robot_v_pref, human_radius,
                                human_v_pref, int(bool(randomize_attributes)), 0)
        robot = torch.empty((count, 9), dtype=F64, device=self.device)
        humans = torch.empty((count, scenarios.MIXED_MAX_HUMANS, 8), dtype=F64, device=self.device)
        counts = torch.empty((count,), dtype=I32, device=self.device)
        seed = scenarios.case_seed(phase, int(first_case))
        check(self.lib.aem_generate_mixed_scenarios(self.h, ctypes.c_uint32(seed & 0xffffffff), int(count), ctypes.byref(p),
                                                    _ptr(robot), _ptr(humans), _ptr(counts), self._stream()))
        return robot, humans, counts

    # ---- value-network regression step (trainer.py:47-58) ------------------------------------------------
    def train_grad(self, params, states, targets, grad, loss=None, values=None, steps=None):
        """forward + MSE + backward of the flag-5 network on [B, n, 13] rotated states (batch-global ACT halting);
        params / grad: flat float32[113752] in blob order; returns (grad, loss[1], values[B], steps[1])"""
        B, n, _ = states.shape
        self._chk(states, F32, (B, n, 13)); self._chk(targets.view(-1), F32, (B,))
        self._chk(params, F32, (_lib.NUM_WEIGHTS,)); self._chk(grad, F32, (_lib.NUM_WEIGHTS,))
        loss = torch.empty((1,), dtype=F32, device=self.device) if loss is None else loss
        values = torch.empty((B,), dtype=F32, device=self.device) if values is None else values
        steps = torch.empty((1,), dtype=I32, device=self.device) if steps is None else steps
        check(self.lib.aem_train_grad(self.h, B, n, _ptr(params), _ptr(states), _ptr(targets), _ptr(grad), _ptr(loss),
                                      _ptr(values), _ptr(steps), self._stream()))
        return grad, loss, values, steps

    def sample_batch(self, states, values, size, k, seed, counter, want_index=False):
        """k of the first `size` (state, value) pairs of a replay ring, uniformly without replacement -> ([k, n, 13], [k, 1])"""
        cap, n, _ = states.shape
        self._chk(states, F32, (cap, n, 13)); self._chk(values, F32, (cap, 1))
        out_s = torch.empty((k, n, 13), dtype=F32, device=self.device)
        out_v = torch.empty((k, 1), dtype=F32, device=self.device)
        idx = torch.empty((k,), dtype=torch.int64, device=self.device) if want_index else None
        check(self.lib.aem_sample_batch(self.h, int(size), int(k), n, _ptr(states), _ptr(values), ctypes.c_uint64(seed & (2 ** 64 - 1)),
                                        ctypes.c_uint64(counter), _ptr(out_s), _ptr(out_v), _ptr(idx), self._stream()))
        return (out_s, out_v, idx) if want_index else (out_s, out_v)

    def adam_step(self, params, grad, exp_avg, exp_avg_sq, lr, betas, eps, step, grad_scale=1.0):
        """torch.optim.Adam.step() on flat float32 vectors (no weight decay / amsgrad); step is 1-based"""
        for t in (params, grad, exp_avg, exp_avg_sq):
            self._chk(t, F32, (params.numel(),))
        check(self.lib.aem_adam_step(self.h, params.numel(), _ptr(params), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq),
                                     float(lr), float(betas[0]), float(betas[1]), float(eps), int(step), float(grad_scale),
                                     self._stream()))

    def train_step_adam(self, ring_states, ring_values, size, k, seed, dev_state, adam_scal, params, grad, exp_avg, exp_avg_sq,
                        lr, betas, eps, batch_states, batch_values, loss, loss_sum, steps):
        """one iteration of Trainer.optimize_batch with Adam, all per-step state on the device (graph-capturable)"""
        cap, n, _ = ring_states.shape
        check(self.lib.aem_train_step_adam(self.h, int(size), int(k), n, _ptr(ring_states), _ptr(ring_values),
                                           ctypes.c_uint64(seed & (2 ** 64 - 1)), _ptr(dev_state), _ptr(adam_scal), _ptr(params),
                                           _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), float(lr), float(betas[0]), float(betas[1]),
                                           float(eps), _ptr(batch_states), _ptr(batch_values), _ptr(loss), _ptr(loss_sum), _ptr(steps),
                                           self._stream()))

    def sgd_step(self, params, grad, momentum_buf, lr, momentum, step, grad_scale=1.0):
        """torch.optim.SGD(momentum).step() on flat float32 vectors; step is 1-based (the first call initialises the buffer)"""
        check(self.lib.aem_sgd_step(self.h, params.numel(), _ptr(params), _ptr(grad), _ptr(momentum_buf), float(lr), float(momentum),
                                    int(step), float(grad_scale), self._stream()))

    def rmsprop_step(self, params, grad, square_avg, lr, alpha, eps, grad_scale=1.0):
        """torch.optim.RMSprop(alpha, eps).step() on flat float32 vectors"""
        check(self.lib.aem_rmsprop_step(self.h, params.numel(), _ptr(params), _ptr(grad), _ptr(square_avg), float(lr), float(alpha),
                                        float(eps), float(grad_scale), self._stream()))

    # ---- host-buffer call (the e2e path) -------------------------------------------------------
    def decide_step_host(self, robot, humans, global_time, gamma_bar, want_values=False):
        """robot [B,9], humans [B,n,8], global_time [B]: contiguous float64 numpy arrays, updated IN PLACE."""
        B, n, _ = humans.shape
        for a in (robot, humans, global_time):
            assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous
        chosen = np.empty(B, dtype=np.int32)
        reward = np.empty(B, dtype=np.float64)
        done = np.empty(B, dtype=np.int32)
        info = np.empty(B, dtype=np.int32)
        values = np.empty((B, self.A), dtype=np.float64) if want_values else None
        vp = values.ctypes.data_as(ctypes.c_void_p) if want_values else ctypes.c_void_p(0)
        check(self.lib.aem_decide_step_host(self.h, B, n, float(gamma_bar), robot.ctypes.data_as(ctypes.c_void_p),
                                            humans.ctypes.data_as(ctypes.c_void_p),
                                            global_time.ctypes.data_as(ctypes.c_void_p),
                                            chosen.ctypes.data_as(ctypes.c_void_p),
                                            reward.ctypes.data_as(ctypes.c_void_p),
                                            done.ctypes.data_as(ctypes.c_void_p),
                                            info.ctypes.data_as(ctypes.c_void_p), vp))
        return chosen, reward, done, info, values

    def launch_count(self):
        return int(self.lib.aem_launch_count(self.h))

    def set_active_counts(self, n_active=None, pool_n_active=None):
        """per-environment human counts [B] int32 ('mixed' scenes); None = uniform batches.  The tensors must stay alive."""
        if n_active is not None:
            self._chk(n_active, I32)
        if pool_n_active is not None:
            self._chk(pool_n_active, I32)
        self._active_counts = (n_active, pool_n_active)
        check(self.lib.aem_set_active_counts(self.h, _ptr(n_active), _ptr(pool_n_active)))

    def set_train_dropout(self, p, seed=0):
        """dropout of the fused training step (nn.TransformerEncoderLayer's four sites); 0 = the eval()-mode network"""
        check(self.lib.aem_set_train_dropout(self.h, float(p), ctypes.c_uint64(int(seed) & (2 ** 64 - 1))))

    def train_scratch_generation(self):
        """changes whenever the library reallocated the training scratch: captured graphs of train_step_adam are stale then"""
        return int(self.lib.aem_train_scratch_generation(self.h))


_ENGINES = {}


def get_engine(device=0, role="default"):
    """Process-wide Engine cache: one aem_handle per (device, role).  Roles keep independent action tables /
    env params apart (e.g. the policy's 81-action table vs. the env's single arbitrary action)."""
    idx = device if isinstance(device, int) else (torch.device(device).index or 0)
    key = (idx, role)
    if key not in _ENGINES:
        _ENGINES[key] = Engine(idx)
    return _ENGINES[key]
