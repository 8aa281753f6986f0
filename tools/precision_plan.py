#!/usr/bin/env python
"""precision_plan.py -- CPU emulation of per-layer FP16 split plans of the lookahead kernel (no GPU needed).

Every GEMM of the value network can run as
    x3 : A_hi W_hi + A_lo W_hi + A_hi W_lo   (the shipped FP16x3 arithmetic, ~22 mantissa bits)
    a2 : A_hi W_hi + A_lo W_hi               (weights rounded to fp16)
    w2 : A_hi W_hi + A_hi W_lo               (activations rounded to fp16)
    x1 : A_hi W_hi                           (single pass)
with FP32 accumulation.  This script evaluates a plan {layer group -> mode} in numpy (products exact in float64,
operands rounded exactly like the kernel does) on (env, action) samples of the bench workload and reports, against the
all-fp32 network: value error relative to max(|V|, 1e-2), ACT step-count flips, argmax flips per environment.

    python tools/precision_plan.py [--envs 256] [--weights file.npy] [--plans name,name,...]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["z1", "z2", "r1", "r2", "q", "in", "out", "f1", "f2", "a0", "a2"]


def h16(x):
    return x.astype(np.float16).astype(np.float64)


def split(x):
    hi = h16(x)
    lo = h16(x - hi)
    return hi, lo


class Net(object):
    def __init__(self, flat):
        from aemcarl_b200.weights import unflatten
        self.w = {k: v.astype(np.float64) for k, v in unflatten(flat).items()}
        self.cache = {}

    def wsplit(self, key):
        if key not in self.cache:
            self.cache[key] = split(self.w[key])
        return self.cache[key]

    def lin(self, a, wkey, bkey, mode):
        """a [.., K] float64 (values are fp32-representable), W [N, K]"""
        W = self.w[wkey]
        b = self.w[bkey]
        if mode == "fp32":
            y = a @ W.T
        else:
            whi, wlo = self.wsplit(wkey)
            ahi, alo = split(a)
            y = ahi @ whi.T
            if mode in ("x3", "a2"):
                y = y + alo @ whi.T
            if mode in ("x3", "w2"):
                y = y + ahi @ wlo.T
        y = y + b
        return y.astype(np.float32).astype(np.float64)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def layer_norm(x, g, b):
    m = x.mean(-1, keepdims=True)
    v = ((x - m) ** 2).mean(-1, keepdims=True)
    return (x - m) / np.sqrt(v + 1e-5) * g + b


def forward(net, states, plan):
    """states [S, n, 13] rotated; returns values [S], steps [S], min halting margin [S]"""
    S, n, _ = states.shape
    P = lambda g: plan.get(g, "fp32")
    x = states.astype(np.float64)
    tok0 = x[:, :1, :].copy()
    tok0[:, 0, 6] = 0; tok0[:, 0, 7] = 0; tok0[:, 0, 8] = x[:, 0, 4]; tok0[:, 0, 9] = x[:, 0, 5]
    tok0[:, 0, 10] = x[:, 0, 3]; tok0[:, 0, 11] = 0; tok0[:, 0, 12] = x[:, 0, 3]
    xs = np.concatenate([tok0, x], axis=1)                  # [S, T, 13]
    T = n + 1
    h = np.zeros((S, T, 50))
    Pacc = np.zeros((S, T))
    A = np.zeros((S, T, 50))
    active = np.ones(S, dtype=bool)
    cnt = np.zeros(S, dtype=np.int64)
    margin = np.full(S, np.inf)
    pre = "in_mlp.rnn_cell."
    for step in range(1, 4):
        hx = np.concatenate([h, xs], axis=-1)
        z = sigmoid(np.maximum(net.lin(np.maximum(net.lin(hx, pre + "convz.0.weight", pre + "convz.0.bias", P("z1")), 0),
                                       pre + "convz.2.weight", pre + "convz.2.bias", P("z2")), 0))
        r = sigmoid(np.maximum(net.lin(np.maximum(net.lin(hx, pre + "convr.0.weight", pre + "convr.0.bias", P("r1")), 0),
                                       pre + "convr.2.weight", pre + "convr.2.bias", P("r2")), 0))
        rhx = np.concatenate([r * h, xs], axis=-1)
        q = np.tanh(net.lin(rhx, pre + "convq.weight", pre + "convq.bias", P("q")))
        h = ((1 - z) * h + z * q).astype(np.float32).astype(np.float64)
        p = sigmoid(h @ net.w["in_mlp.ponder_linear.weight"][0] + net.w["in_mlp.ponder_linear.bias"][0])
        Pacc[active] += p[active]
        A[active] += p[active][..., None] * h[active]
        cnt[active] = step
        margin[active] = np.minimum(margin[active], np.abs(Pacc[active] - 0.95).min(axis=1))
        need = (Pacc < 0.95).any(axis=1)
        active = active & need
        if not active.any():
            break
    e = (A / cnt[:, None, None]).astype(np.float32).astype(np.float64)
    for l in range(3):
        pl = "transformer_encoder.layers.%d." % l
        qkv = net.lin(e, pl + "self_attn.in_proj_weight", pl + "self_attn.in_proj_bias", P("in"))
        qh = qkv[..., 0:50].reshape(S, T, 2, 25)
        kh = qkv[..., 50:100].reshape(S, T, 2, 25)
        vh = qkv[..., 100:150].reshape(S, T, 2, 25)
        sc = np.einsum("sihd,sjhd->shij", qh, kh) / 5.0
        sc = sc - sc.max(-1, keepdims=True)
        w = np.exp(sc)
        w = w / w.sum(-1, keepdims=True)
        o = np.einsum("shij,sjhd->sihd", w, vh).reshape(S, T, 50).astype(np.float32).astype(np.float64)
        att = net.lin(o, pl + "self_attn.out_proj.weight", pl + "self_attn.out_proj.bias", P("out"))
        e = layer_norm(e + att, net.w[pl + "norm1.weight"], net.w[pl + "norm1.bias"]).astype(np.float32).astype(np.float64)
        f = np.maximum(net.lin(e, pl + "linear1.weight", pl + "linear1.bias", P("f1")), 0)
        f = net.lin(f, pl + "linear2.weight", pl + "linear2.bias", P("f2"))
        e = layer_norm(e + f, net.w[pl + "norm2.weight"], net.w[pl + "norm2.bias"]).astype(np.float32).astype(np.float64)
    joint = np.concatenate([xs[:, 0, :6], e[:, 0, :]], axis=-1)
    a = np.maximum(net.lin(joint, "action_mlp.0.weight", "action_mlp.0.bias", P("a0")), 0)
    a = np.maximum(net.lin(a, "action_mlp.2.weight", "action_mlp.2.bias", P("a2")), 0)
    v = a @ net.w["action_mlp.4.weight"][0] + net.w["action_mlp.4.bias"][0]
    return v, cnt, margin


def np_rotate(in14):
    """cadrl.py:239-274 vectorised, float32 like the reference"""
    s = in14.astype(np.float32)
    dx, dy = s[:, 5] - s[:, 0], s[:, 6] - s[:, 1]
    rot = np.arctan2(dy, dx)
    c, sn = np.cos(rot), np.sin(rot)
    o = np.zeros((s.shape[0], 13), dtype=np.float32)
    o[:, 0] = np.sqrt(dx * dx + dy * dy)
    o[:, 1] = s[:, 7]
    o[:, 3] = s[:, 4]
    o[:, 4] = s[:, 2] * c + s[:, 3] * sn
    o[:, 5] = s[:, 3] * c - s[:, 2] * sn
    rx, ry = s[:, 9] - s[:, 0], s[:, 10] - s[:, 1]
    o[:, 6] = rx * c + ry * sn
    o[:, 7] = ry * c - rx * sn
    o[:, 8] = s[:, 11] * c + s[:, 12] * sn
    o[:, 9] = s[:, 12] * c - s[:, 11] * sn
    o[:, 10] = s[:, 13]
    o[:, 11] = np.sqrt((s[:, 0] - s[:, 9]) ** 2 + (s[:, 1] - s[:, 10]) ** 2)
    o[:, 12] = s[:, 4] + s[:, 13]
    return o


def workload(B, n, warm_steps=12, seed=0):
    """B environment states of the bench stream advanced a few ORCA steps -> states [B*81, n, 13], rewards [B, 81]"""
    from aemcarl_b200 import scenarios
    from oracle import pyoracle as po
    params = po.default_params()
    actions = po.build_action_space()
    robot, humans = scenarios.generate_pool("train", 0, B, n, "circle_crossing")
    gt = np.zeros(B)
    rs = np.random.RandomState(seed)
    for it in range(warm_steps):
        nv = po.orca(humans, robot, params, 8)
        chosen = rs.randint(0, 81, size=B).astype(np.int32)
        po.env_apply(robot, humans, nv, gt, actions, chosen, params)
    nv = po.orca(humans, robot, params, 8)
    rew, info, dmin, hnext = po.env_lookahead(robot, humans, nv, gt, actions, params, 8)
    A = actions.shape[0]
    in14 = np.zeros((B, A, n, 14), dtype=np.float32)
    in14[:, :, :, 0] = (robot[:, None, 0] + actions[None, :, 0] * params.time_step)[:, :, None]
    in14[:, :, :, 1] = (robot[:, None, 1] + actions[None, :, 1] * params.time_step)[:, :, None]
    in14[:, :, :, 2] = actions[None, :, 0, None]
    in14[:, :, :, 3] = actions[None, :, 1, None]
    in14[:, :, :, 4:9] = robot[:, None, None, 4:9]
    in14[:, :, :, 9:14] = hnext[:, None, :, :]
    st = np_rotate(in14.reshape(-1, 14)).reshape(B * A, n, 13)
    return st, rew, params


PLANS = {
    "x3": {g: "x3" for g in GROUPS},
    "x1": {g: "x1" for g in GROUPS},
    "a2": {g: "a2" for g in GROUPS},
    "w2": {g: "w2" for g in GROUPS},
    "gru3_enc1": dict({g: "x3" for g in ["z1", "z2", "r1", "r2", "q"]}, **{g: "x1" for g in ["in", "out", "f1", "f2", "a0", "a2"]}),
    "gru3_enc_a2": dict({g: "x3" for g in ["z1", "z2", "r1", "r2", "q"]}, **{g: "a2" for g in ["in", "out", "f1", "f2", "a0", "a2"]}),
    "gru3_enc_w2": dict({g: "x3" for g in ["z1", "z2", "r1", "r2", "q"]}, **{g: "w2" for g in ["in", "out", "f1", "f2", "a0", "a2"]}),
}
for g in GROUPS:                       # one group degraded at a time
    for m in ("x1", "a2", "w2"):
        PLANS["only_%s_%s" % (g, m)] = dict({k: "x3" for k in GROUPS}, **{g: m})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=128)
    ap.add_argument("--humans", type=int, default=5)
    ap.add_argument("--weights", default=None)
    ap.add_argument("--plans", default="x3,x1,a2,w2,gru3_enc1,gru3_enc_a2,gru3_enc_w2")
    args = ap.parse_args()
    from aemcarl_b200 import weights as wts
    from oracle import pyoracle as po
    W = wts.synthetic_weights(0) if not args.weights else np.load(args.weights).astype(np.float32)
    st, rew, params = workload(args.envs, args.humans)
    net = Net(W)
    v0, k0, m0 = forward(net, st, {})
    vo, ko = po.value_forward_batch(W, st, nthreads=8)
    print("numpy fp32 net vs C oracle: value err %.2e, step mismatches %d of %d; ACT hist %s" % (
        np.max(np.abs(v0 - vo) / np.maximum(np.abs(vo), 1e-2)), int((k0 != ko).sum()), len(ko), np.bincount(k0, minlength=4).tolist()))
    gbar = pow(0.9, params.time_step)
    B = args.envs
    tot0 = rew + gbar * v0.reshape(B, 81)
    best0 = tot0.argmax(1)
    names = args.plans.split(",")
    if names == ["all"]:
        names = list(PLANS)
    print("%-16s %10s %10s %10s %8s %8s %s" % ("plan", "max_err", "p99.9_err", "rms_err", "K_flips", "argflips", "worst near-tie gap of a flip"))
    for name in names:
        v, k, m = forward(net, st, PLANS[name])
        same = k == k0
        err = np.abs(v - v0) / np.maximum(np.abs(v0), 1e-2)
        e = err[same]
        tot = rew + gbar * v.reshape(B, 81)
        best = tot.argmax(1)
        fl = np.nonzero(best != best0)[0]
        gap = [abs(tot0[b, best0[b]] - tot0[b, best[b]]) / max(abs(tot0[b, best0[b]]), 1e-2) for b in fl]
        print("%-16s %10.2e %10.2e %10.2e %8d %8d %s" % (name, e.max(), np.quantile(e, 0.999), np.sqrt((e ** 2).mean()),
                                                          int((~same).sum()), len(fl), "%.1e" % max(gap) if gap else "-"))


if __name__ == "__main__":
    main()
