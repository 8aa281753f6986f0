#!/usr/bin/env python
"""h16_dev.py -- development harness of the lookahead kernels (GPU).

    python tools/h16_dev.py check [precision]   values / ACT steps against the FP32 kernel and the CPU oracle
    python tools/h16_dev.py time  [precision]   CUDA-event timing of aem_lookahead at the bench size (L2 flushed)
    python tools/h16_dev.py once  [precision]   two launches (timeline / watchdog builds: AEM_LIB_PATH, AEM_H16_TIMELINE)

Environment: AEM_LIB_PATH (experiment build of the library), AEM_H16_OPT (kernel option bits).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aemcarl_b200 import scenarios, weights as wts                     # noqa: E402
from aemcarl_b200.engine import Engine, make_env_params                # noqa: E402
from aemcarl_b200.policy_math import build_action_space                # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def states(B, n, seed=0):
    robot, humans = scenarios.generate_pool("train", 0, B, n, "circle_crossing")
    rng = np.random.RandomState(seed)
    humans[:, :, 2:4] = rng.uniform(-0.7, 0.7, (B, n, 2))
    robot[:, 0:2] += rng.uniform(-1, 1, (B, 2))
    hnext = np.concatenate([humans[:, :, 0:2] + 0.2 * humans[:, :, 2:4], humans[:, :, 2:5]], axis=2)
    return robot, hnext


def main():
    cmd = sys.argv[1]
    prec = sys.argv[2] if len(sys.argv) > 2 else "fp16x3"
    eng = Engine(0)
    actions = build_action_space(1.0)
    eng.set_action_space(actions)
    eng.set_env_params(make_env_params())
    tag = "lib=%s opt=%s prec=%s" % (os.path.basename(os.environ.get("AEM_LIB_PATH", "default")),
                                    os.environ.get("AEM_H16_OPT", "-"), prec)
    if cmd == "diag5":         # h after 2 plain GRU steps with parts of the cell switched off
        from oracle import pyoracle as po
        rng = np.random.RandomState(3)
        offs = wts.weight_offsets()
        T, S = 6, 77
        x = rng.randn(S * T, 13).astype(np.float32)
        eng.set_precision(prec)

        def zero_h(W, key, rows, only=None):
            o, sh = offs[key]
            M = W[o:o + int(np.prod(sh))].reshape(sh)
            if only is None:
                M[:, :50] = 0.0
            else:
                keep = M[:, only].copy()
                M[:, :50] = 0.0
                M[:, only] = keep
        for name in ("full", "q_noh", "q_noh_z_noh", "q_noh_z_noh_r", "z_noh", "r_noh", "q_only_h0", "q_only_h24", "q_only_h25", "q_only_h49",
                     "z_only_h0", "z_only_h24"):
            W = wts.synthetic_weights(0).copy()
            if name.startswith("q_noh"):
                zero_h(W, "in_mlp.rnn_cell.convq.weight", 50)
            if "z_noh" in name:
                zero_h(W, "in_mlp.rnn_cell.convz.0.weight", 100)
            if name.endswith("_r") or name == "r_noh":
                zero_h(W, "in_mlp.rnn_cell.convr.0.weight", 100)
            if name.startswith("q_only_h"):
                zero_h(W, "in_mlp.rnn_cell.convq.weight", 50, only=int(name[8:]))
            if name.startswith("z_only_h"):
                zero_h(W, "in_mlp.rnn_cell.convz.0.weight", 100, only=int(name[8:]))
                zero_h(W, "in_mlp.rnn_cell.convq.weight", 50)
            eng.load_weights(W)
            ro, _ = po.gru_act(W, x, T, act_fixed=True, act_steps=2)
            eng.set_act_mode(True, 2)
            o, st = eng.gru_act(dev(x), T)
            err = np.abs(o.cpu().numpy() - ro)
            print("DIAG5 %s [%s]: h2 max err %.2e" % (tag, name, err.max()), flush=True)
        eng.set_act_mode(False, 1)
        return
    if cmd == "diag4":         # plain GRU steps (act_fixed): is h_k right?
        from oracle import pyoracle as po
        rng = np.random.RandomState(3)
        W = wts.synthetic_weights(0)
        eng.load_weights(W)
        T, S = 6, 77
        x = rng.randn(S * T, 13).astype(np.float32)
        eng.set_precision(prec)
        for k in (1, 2, 3):
            ro, _ = po.gru_act(W, x, T, act_fixed=True, act_steps=k)
            eng.set_act_mode(True, k)
            o, st = eng.gru_act(dev(x), T)
            err = np.abs(o.cpu().numpy() - ro).reshape(S, T, 50)
            print("DIAG4 %s: h after %d plain GRU steps: max err %.2e; per column max: %s" % (
                tag, k, err.max(), " ".join("%.0e" % e for e in err.max(axis=(0, 1)))), flush=True)
        eng.set_act_mode(False, 1)
        return
    if cmd == "diag3":         # GRU-ACT alone: where are the errors (sample K, token, column, tile position)?
        from oracle import pyoracle as po
        rng = np.random.RandomState(3)
        W = wts.synthetic_weights(0)
        eng.load_weights(W)
        T = 6
        S = 77
        x = rng.randn(S * T, 13).astype(np.float32)
        ro, rsteps = po.gru_act(W, x, T)
        eng.set_precision(prec)
        o, st = eng.gru_act(dev(x), T)
        o, st = o.cpu().numpy(), st.cpu().numpy()
        err = np.abs(o - ro).reshape(S, T, 50)
        print("DIAG3 %s: max %.2e" % (tag, err.max()))
        for k in (1, 2, 3):
            m = rsteps == k
            if m.any():
                print("  K=%d samples %d: max err %.2e median of sample max %.2e" % (k, m.sum(), err[m].max(), np.median(err[m].max(axis=(1, 2)))))
        print("  per token max:", ["%.1e" % e for e in err.max(axis=(0, 2))])
        print("  per column max:", ["%.0e" % e for e in err.max(axis=(0, 1))])
        print("  per sample max (first 42 = two tiles):", ["%.0e" % e for e in err.max(axis=(1, 2))[:42]])
        print("  steps (first 42):", rsteps[:42].tolist())
        ratio = (o.reshape(S, T, 50) / np.where(np.abs(ro) > 1e-3, ro, np.nan).reshape(S, T, 50))
        print("  out/ref ratio per sample (median over elements, first 42):", ["%.3f" % r for r in np.nanmedian(ratio, axis=(1, 2))[:42]])
        return
    if cmd == "diag2":         # which phase is wrong?  GRU-ACT alone, then weight variants that neutralise one component
        from oracle import pyoracle as po
        rng = np.random.RandomState(3)
        offs = wts.weight_offsets()

        def variant(name):
            W = wts.synthetic_weights(0).copy()
            if name == "ln_identity":
                for k, (o, sh) in offs.items():
                    if "norm" in k:
                        W[o:o + int(np.prod(sh))] = 1.0 if k.endswith("weight") else 0.0
            if name == "k1":
                o, sh = offs["in_mlp.ponder_linear.bias"]; W[o] = 10.0
            if name == "k3":
                o, sh = offs["in_mlp.ponder_linear.bias"]; W[o] = -10.0
            return W
        for name in ("normal", "ln_identity", "k1", "k3"):
            W = variant(name)
            eng.load_weights(W)
            T = 6
            x = rng.randn(77 * T, 13).astype(np.float32)
            ro, rsteps = po.gru_act(W, x, T)
            eng.set_precision(prec)
            o, st = eng.gru_act(dev(x), T)
            o, st = o.cpu().numpy(), st.cpu().numpy()
            B, n = 64, 5
            robot, hnext = states(B, n, seed=B + n)
            args = (dev(robot), dev(hnext), dev(np.zeros((B, 81))), 1.0)
            eng.set_precision("fp32")
            _, vr, _, sr = [t.cpu().numpy().ravel() for t in eng.lookahead(*args)]
            eng.set_precision(prec)
            _, v, _, s = [t.cpu().numpy().ravel() for t in eng.lookahead(*args)]
            rel = np.abs(v - vr) / np.maximum(np.abs(vr), 1e-2)
            print("DIAG2 %s [%s]: gru_act max abs err %.2e (steps equal: %s, hist %s); lookahead median rel %.2e max %.2e flips %d"
                  % (tag, name, np.abs(o - ro).max(), bool((st == rsteps).all()), np.bincount(rsteps, minlength=4).tolist(),
                     np.median(rel), rel.max(), int((s != sr).sum())), flush=True)
        return
    if cmd == "diag":          # where in a tile do the bad samples sit?
        W = wts.synthetic_weights(0)
        eng.load_weights(W)
        B, n = 300, 5
        robot, hnext = states(B, n, seed=B + n)
        args = (dev(robot), dev(hnext), dev(np.zeros((B, 81))), 1.0)
        eng.set_precision("fp32")
        _, vr, _, sr = [t.cpu().numpy().ravel() for t in eng.lookahead(*args)]
        eng.set_precision(prec)
        _, v, _, s = [t.cpu().numpy().ravel() for t in eng.lookahead(*args)]
        rel = np.abs(v - vr) / np.maximum(np.abs(vr), 1e-2)
        S = 21
        bad = np.nonzero((rel > 1e-4) & (s == sr))[0]
        flip = np.nonzero(s != sr)[0]
        print("DIAG %s: bad %d flips %d of %d; max rel %.2e; median rel of bad %.2e" % (tag, len(bad), len(flip), v.size, rel.max(),
              np.median(rel[bad]) if len(bad) else 0.0))
        print("  bad pos in tile:", np.bincount(bad % S, minlength=S).tolist())
        print("  flip pos in tile:", np.bincount(flip % S, minlength=S).tolist())
        print("  flips mine->ref:", sorted(set(zip(s[flip].tolist(), sr[flip].tolist()))), " K hist mine", np.bincount(s, minlength=4).tolist(),
              "ref", np.bincount(sr, minlength=4).tolist())
        print("  rel quantiles (all):", ["%.1e" % q for q in np.quantile(rel, [0.5, 0.9, 0.99, 0.999])])
        tiles = np.unique(bad // S)
        print("  tiles with a bad sample: %d of %d; bad per such tile %.1f" % (len(tiles), (v.size + S - 1) // S, len(bad) / max(1, len(tiles))))
        return
    if cmd == "check":
        from oracle import pyoracle as po
        worst = 0.0
        for wseed, pb, wfile in ((0, None, None), (7, -0.05, None), (0, None, "gpurun_out/trained_flat.npy")):
            if wfile and not os.path.exists(os.path.join(ROOT, wfile)):
                continue
            W = np.load(os.path.join(ROOT, wfile)).astype(np.float32) if wfile else wts.synthetic_weights(wseed, pb)
            eng.load_weights(W)
            for B, n in ((1, 5), (64, 5), (4096, 5), (33, 1), (40, 10), (20, 18), (6, 32)):
                robot, hnext = states(B, n, seed=B + n)
                rewards = np.zeros((B, 81))
                args = (dev(robot), dev(hnext), dev(rewards), 1.0)
                eng.set_precision("fp32")
                _, vr, _, sr = [t.cpu().numpy() for t in eng.lookahead(*args)]
                eng.set_precision(prec)
                _, v, _, s = [t.cpu().numpy() for t in eng.lookahead(*args)]
                flips = int((s != sr).sum())
                ok = s == sr
                rel = np.abs(v - vr) / np.maximum(np.abs(vr), 1e-2)
                line = "w=%s B=%d n=%d: max rel %.2e (same K), K flips %d of %d, hist %s, nan %d" % (
                    wfile or (wseed, pb), B, n, rel[ok].max(), flips, s.size, np.bincount(s.ravel(), minlength=4).tolist(),
                    int(np.isnan(v).sum()))
                if B <= 8:
                    rv, rnet, rbest, rsteps = po.lookahead(W, robot, hnext, rewards, actions, gamma_bar=1.0, nthreads=8)
                    relo = np.abs(v - rnet) / np.maximum(np.abs(rnet), 1e-2)
                    line += "; vs oracle %.2e, K flips %d" % (relo[s == rsteps].max(), int((s != rsteps).sum()))
                print(line, flush=True)
                worst = max(worst, float(rel[ok].max()))
        print("CHECK %s worst %.2e %s" % (tag, worst, "OK" if worst <= 1e-4 else "FAIL"), flush=True)
        return
    W = wts.synthetic_weights(0)
    eng.load_weights(W)
    eng.set_precision(prec)
    B, n = int(os.environ.get("DEV_ENVS", "4096")), int(os.environ.get("DEV_HUMANS", "5"))
    robot, hnext = states(B, n)
    args = (dev(robot), dev(hnext), dev(np.zeros((B, 81))), 1.0)
    out = eng.lookahead(*args)
    torch.cuda.synchronize()
    if cmd == "once":
        eng.lookahead(*args, out=out)
        torch.cuda.synchronize()
        print("ONCE %s done" % tag)
        return
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    reps = 20
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for _ in range(3):
        eng.lookahead(*args, out=out)
    for a, b in ev:
        flush.zero_()
        a.record()
        eng.lookahead(*args, out=out)
        b.record()
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    print("TIME %s B=%d n=%d: lookahead %.3f ms (min %.3f) -> %.0f k env-steps/s kernel-only" % (
        tag, B, n, ms.mean(), ms.min(), B / ms.mean()), flush=True)


if __name__ == "__main__":
    main()
