"""ctypes front-end of the CPU oracle (oracle/aem_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs.  The product package (aemcarl_b200) never imports it.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libaem_oracle.so")

NUM_WEIGHTS = 113752
NUM_ACTIONS = 81

INFO_NOTHING, INFO_DANGER, INFO_REACHGOAL, INFO_COLLISION, INFO_TIMEOUT = range(5)


class EnvParams(ctypes.Structure):
    _fields_ = [
        ("time_step", ctypes.c_double),
        ("time_limit", ctypes.c_double),
        ("success_reward", ctypes.c_double),
        ("collision_penalty", ctypes.c_double),
        ("discomfort_dist", ctypes.c_double),
        ("agent_timestep", ctypes.c_double),
        ("human_timestep", ctypes.c_double),
        ("reward_increment", ctypes.c_double),
        ("position_variance", ctypes.c_double),
        ("direction_variance", ctypes.c_double),
        ("robot_visible", ctypes.c_int32),
        ("pad_", ctypes.c_int32),
        ("orca_safety_space", ctypes.c_double),
    ]


def default_params(**kw):
    """env.config defaults with the README reward flags (README.md:56)."""
    p = EnvParams(0.2, 20.0, 1.0, -0.25, 0.2, 0.4, 0.5, 2.0, 2.0, 2.0, 0, 0, 0.0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only; no reference sources involved)."""
    src = os.path.join(_HERE, "aem_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "aem_oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "CC=gcc", "-B", "libaem_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.aemo_occupancy.restype = ctypes.c_double
        _lib.aemo_num_weights.restype = ctypes.c_int
        _lib.aemo_act_margin.restype = ctypes.c_float
        assert _lib.aemo_num_weights() == NUM_WEIGHTS
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


F, D, I32, U8 = ctypes.c_float, ctypes.c_double, ctypes.c_int32, ctypes.c_uint8


def build_action_space(v_pref=1.0, speed_samples=5, rotation_samples=16):
    out = np.zeros((speed_samples * rotation_samples + 1, 2), dtype=np.float64)
    lib().aemo_build_action_space(D(v_pref), speed_samples, rotation_samples, _p(out, D))
    return out


def rotate(in14):
    in14 = _f32(in14).reshape(-1, 14)
    out = np.zeros((in14.shape[0], 13), dtype=np.float32)
    for i in range(in14.shape[0]):
        lib().aemo_rotate(_p(in14[i], F), _p(out[i], F))
    return out


def value_forward(W, state, act_fixed=False, act_steps=1):
    W = _f32(W)
    state = _f32(state)
    n = state.shape[0]
    v = F(0)
    s = ctypes.c_int(0)
    lib().aemo_value_forward(_p(W, F), _p(state, F), n, int(act_fixed), int(act_steps), ctypes.byref(v),
                             ctypes.byref(s))
    return v.value, s.value


def act_margin(W, state):
    """min |accum_p - 0.95| over tokens and executed ACT steps for one sample (state [n][13], rotated)"""
    W = _f32(W)
    state = _f32(state)
    return float(lib().aemo_act_margin(_p(W, F), _p(state, F), state.shape[0]))


def rotated_state(robot9, humans_next, action, dt=0.2):
    """the [n][13] network input of one (env, action) sample exactly as aemo_lookahead builds it"""
    r = np.asarray(robot9, dtype=np.float64)
    hn = np.asarray(humans_next, dtype=np.float64).reshape(-1, 5)
    in14 = np.zeros((hn.shape[0], 14), dtype=np.float32)
    in14[:, 0] = np.float32(r[0] + action[0] * dt)
    in14[:, 1] = np.float32(r[1] + action[1] * dt)
    in14[:, 2], in14[:, 3] = np.float32(action[0]), np.float32(action[1])
    in14[:, 4:9] = r[4:9].astype(np.float32)
    in14[:, 9:14] = hn.astype(np.float32)
    return rotate(in14)


def value_forward_batch(W, states, act_fixed=False, act_steps=1, nthreads=1):
    W = _f32(W)
    states = _f32(states)
    S, n, _ = states.shape
    values = np.zeros(S, dtype=np.float32)
    steps = np.zeros(S, dtype=np.uint8)
    lib().aemo_value_forward_batch(_p(W, F), S, n, _p(states, F), int(act_fixed), int(act_steps), _p(values, F),
                                   _p(steps, U8), nthreads)
    return values, steps


def gru_act(W, x, T, act_fixed=False, act_steps=1):
    W = _f32(W)
    x = _f32(x).reshape(-1, 13)
    S = x.shape[0] // T
    out = np.zeros((S * T, 50), dtype=np.float32)
    steps = np.zeros(S, dtype=np.uint8)
    lib().aemo_gru_act(_p(W, F), S, T, _p(x, F), int(act_fixed), int(act_steps), _p(out, F), _p(steps, U8))
    return out, steps


def lookahead(W, robot, humans_next, rewards, actions, dt=0.2, gamma_bar=None, act_fixed=False, act_steps=1,
              nthreads=1):
    W = _f32(W)
    robot = _f64(robot).reshape(-1, 9)
    B = robot.shape[0]
    humans_next = _f64(humans_next).reshape(B, -1, 5)
    n = humans_next.shape[1]
    actions = _f64(actions).reshape(-1, 2)
    A = actions.shape[0]
    rewards = _f64(rewards).reshape(B, A)
    if gamma_bar is None:
        gamma_bar = pow(0.9, dt * 1.0)
    values = np.zeros((B, A), dtype=np.float64)
    net = np.zeros((B, A), dtype=np.float32)
    best = np.zeros(B, dtype=np.int32)
    steps = np.zeros((B, A), dtype=np.uint8)
    lib().aemo_lookahead(_p(W, F), B, n, A, _p(robot, D), _p(humans_next, D), _p(rewards, D), _p(actions, D),
                         D(dt), D(gamma_bar), int(act_fixed), int(act_steps), _p(values, D), _p(net, F),
                         _p(best, I32), _p(steps, U8), nthreads)
    return values, net, best, steps


def orca(humans, robot, params, nthreads=1):
    humans = _f64(humans)
    B, n, _ = humans.shape
    robot = _f64(robot).reshape(B, 9)
    out = np.zeros((B, n, 2), dtype=np.float64)
    lib().aemo_orca(B, n, _p(humans, D), _p(robot, D), ctypes.byref(params), _p(out, D), nthreads)
    return out


def orca_agent(pos, vel, radius, max_speed, pref_vel, others, neighbor_dist=10.0, max_neighbors=10,
               time_horizon=5.0, time_step=0.2, want_lines=False):
    pos, vel, pref = _f32(pos), _f32(vel), _f32(pref_vel)
    others = _f32(others).reshape(-1, 5)
    m = others.shape[0]
    nv = np.zeros(2, dtype=np.float32)
    nb = np.zeros(max(m, 1), dtype=np.int32)
    cnt = ctypes.c_int(0)
    lines = np.zeros((max(m, 1), 4), dtype=np.float32)
    lib().aemo_orca_agent(_p(pos, F), _p(vel, F), F(radius), F(max_speed), _p(pref, F), m, _p(others, F),
                          F(neighbor_dist), int(max_neighbors), F(time_horizon), F(time_step), _p(nv, F),
                          _p(nb, I32), ctypes.byref(cnt), _p(lines, F))
    if want_lines:
        return nv, nb[:cnt.value].copy(), lines[:cnt.value].copy()
    return nv


def occupancy(hum5, rx, ry, robot_radius=0.3, human_timestep=0.5, pos_var=2.0, dir_var=2.0):
    hum5 = _f64(hum5).reshape(-1, 5)
    return lib().aemo_occupancy(_p(hum5, D), hum5.shape[0], D(rx), D(ry), D(robot_radius), D(human_timestep),
                                D(pos_var), D(dir_var))


def env_lookahead(robot, humans, new_vel, global_time, actions, params, nthreads=1):
    humans = _f64(humans)
    B, n, _ = humans.shape
    robot = _f64(robot).reshape(B, 9)
    new_vel = _f64(new_vel).reshape(B, n, 2)
    global_time = _f64(global_time).reshape(B)
    actions = _f64(actions).reshape(-1, 2)
    A = actions.shape[0]
    rewards = np.zeros((B, A), dtype=np.float64)
    info = np.zeros((B, A), dtype=np.int32)
    dmin = np.zeros((B, A), dtype=np.float64)
    hnext = np.zeros((B, n, 5), dtype=np.float64)
    lib().aemo_env_lookahead(B, n, A, _p(robot, D), _p(humans, D), _p(new_vel, D), _p(global_time, D),
                             _p(actions, D), ctypes.byref(params), _p(rewards, D), _p(info, I32), _p(dmin, D),
                             _p(hnext, D), nthreads)
    return rewards, info, dmin, hnext


def env_apply(robot, humans, new_vel, global_time, actions, chosen, params):
    """In place on robot/humans/global_time (must be contiguous float64 arrays)."""
    B, n, _ = humans.shape
    assert robot.dtype == np.float64 and humans.dtype == np.float64 and global_time.dtype == np.float64
    assert robot.flags.c_contiguous and humans.flags.c_contiguous and global_time.flags.c_contiguous
    new_vel = _f64(new_vel)
    actions = _f64(actions).reshape(-1, 2)
    chosen = np.ascontiguousarray(chosen, dtype=np.int32)
    lib().aemo_env_apply(B, n, actions.shape[0], _p(robot, D), _p(humans, D), _p(new_vel, D), _p(global_time, D),
                         _p(actions, D), _p(chosen, I32), ctypes.byref(params))


# ----------------------------------------------------------------------------------------------
# state_dict <-> flat blob (order documented in aem_oracle.h / include/aemcarl_b200.h)
# ----------------------------------------------------------------------------------------------
def weight_keys():
    keys = []
    for g in ("convz", "convr"):
        for i in (0, 2):
            keys += ["in_mlp.rnn_cell.%s.%d.weight" % (g, i), "in_mlp.rnn_cell.%s.%d.bias" % (g, i)]
    keys += ["in_mlp.rnn_cell.convq.weight", "in_mlp.rnn_cell.convq.bias",
             "in_mlp.ponder_linear.weight", "in_mlp.ponder_linear.bias"]
    for l in range(3):
        pre = "transformer_encoder.layers.%d." % l
        keys += [pre + k for k in ("self_attn.in_proj_weight", "self_attn.in_proj_bias",
                                   "self_attn.out_proj.weight", "self_attn.out_proj.bias",
                                   "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
                                   "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias")]
    for i in (0, 2, 4):
        keys += ["action_mlp.%d.weight" % i, "action_mlp.%d.bias" % i]
    return keys


def flatten_state_dict(sd):
    parts = [np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k],
                        dtype=np.float32).reshape(-1) for k in weight_keys()]
    flat = np.concatenate(parts)
    assert flat.size == NUM_WEIGHTS, flat.size
    return flat
