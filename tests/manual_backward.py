"""Hand-derived forward / backward of the flag-5 value network (qnetwork_factory.py:641-687, components.py:38-137) in
numpy float64, sample by sample and layer by layer -- the derivation `csrc/train.cu` follows, written out so that it can
be checked against torch autograd on the CPU (tests/test_host_logic.py) before the CUDA kernel is checked against
autograd on the GPU.  Test infrastructure only."""
import numpy as np

from aemcarl_b200 import weights as wts


def _sig(a):
    return 1.0 / (1.0 + np.exp(-a))


def _ln_fwd(s, g, b, eps=1e-5):
    mu = s.mean(axis=1, keepdims=True)
    var = ((s - mu) ** 2).mean(axis=1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (s - mu) * rstd
    return xhat * g + b, xhat, rstd


def _ln_bwd(dy, xhat, rstd, g):
    dxhat = dy * g
    m1 = dxhat.mean(axis=1, keepdims=True)
    m2 = (dxhat * xhat).mean(axis=1, keepdims=True)
    return rstd * (dxhat - m1 - xhat * m2), (dy * xhat).sum(0), dy.sum(0)


def gru_probe(W, x, K_max=3):
    """accumulated halting probability after each step, per token row: what decides the batch-global step count"""
    h = np.zeros((x.shape[0], 50))
    P = np.zeros(x.shape[0])
    out = []
    for _ in range(K_max):
        h, _ = _gru_step(W, x, h)
        P = P + _sig(h @ W["in_mlp.ponder_linear.weight"][0] + W["in_mlp.ponder_linear.bias"][0])
        out.append(P.copy())
    return out


def _gru_step(W, x, h):
    c = "in_mlp.rnn_cell."
    hx = np.concatenate([h, x], axis=1)
    az1 = np.maximum(hx @ W[c + "convz.0.weight"].T + W[c + "convz.0.bias"], 0)
    az2 = np.maximum(az1 @ W[c + "convz.2.weight"].T + W[c + "convz.2.bias"], 0)
    ar1 = np.maximum(hx @ W[c + "convr.0.weight"].T + W[c + "convr.0.bias"], 0)
    ar2 = np.maximum(ar1 @ W[c + "convr.2.weight"].T + W[c + "convr.2.bias"], 0)
    z, r = _sig(az2), _sig(ar2)
    rhx = np.concatenate([r * h, x], axis=1)
    q = np.tanh(rhx @ W[c + "convq.weight"].T + W[c + "convq.bias"])
    hn = (1 - z) * h + z * q
    return hn, dict(h=h, hx=hx, az1=az1, az2=az2, ar1=ar1, ar2=ar2, z=z, r=r, rhx=rhx, q=q, hn=hn)


def tokens_of(state):
    """[n,13] -> [n+1,13] with the robot pseudo-token first (qnetwork_factory.py:651-663)"""
    tok = state[0].copy()
    tok[6] = tok[7] = tok[11] = 0.0
    tok[8], tok[9], tok[10], tok[12] = state[0, 4], state[0, 5], state[0, 3], state[0, 3]
    return np.concatenate([tok[None], state], axis=0)


def batch_steps(flat, states, K_max=3, eps=0.05):
    """components.py:132-134: the step count is decided by the whole batch"""
    W = {k: v.astype(np.float64) for k, v in wts.unflatten(flat).items()}
    K = K_max
    probes = [gru_probe(W, tokens_of(s.astype(np.float64)), K_max) for s in states]
    for k in range(K_max):
        if not any((p[k] < 1 - eps).any() for p in probes):
            K = k + 1
            break
    return K


def sample_forward_backward(flat, state, target, K, batch, masks=None):
    """one sample: value, and d(mean squared error over `batch` samples)/d(parameters) as a flat float64 vector.
    masks (optional): per encoder layer a dict of the four inverted-dropout masks of nn.TransformerEncoderLayer, already scaled
    by 1 / (1 - p): "attn" [2, T, T] on the softmax output, "ao" [T, 50] on the attention branch before the residual, "f1"
    [T, 150] after the FFN's ReLU, "f2" [T, 50] on the FFN branch before the residual (the derivation a dropout-active fused
    training step would follow)."""
    W = {k: v.astype(np.float64) for k, v in wts.unflatten(flat).items()}
    G = {k: np.zeros_like(v) for k, v in W.items()}
    state = state.astype(np.float64)
    x = tokens_of(state)
    T = x.shape[0]
    # ---- forward ----
    h = np.zeros((T, 50))
    A = np.zeros((T, 50))
    steps = []
    wp, bp = W["in_mlp.ponder_linear.weight"][0], W["in_mlp.ponder_linear.bias"][0]
    for _ in range(K):
        h, st = _gru_step(W, x, h)
        st["p"] = _sig(h @ wp + bp)
        A = A + st["p"][:, None] * h
        steps.append(st)
    X = A / K
    layers = []
    for l in range(3):
        p = "transformer_encoder.layers.%d." % l
        qkv = X @ W[p + "self_attn.in_proj_weight"].T + W[p + "self_attn.in_proj_bias"]
        q, k, v = qkv[:, :50], qkv[:, 50:100], qkv[:, 100:]
        ctx = np.zeros((T, 50))
        P = np.zeros((2, T, T))
        for hd in range(2):
            sl = slice(25 * hd, 25 * hd + 25)
            s = (q[:, sl] @ k[:, sl].T) / 5.0
            s = s - s.max(axis=1, keepdims=True)
            e = np.exp(s)
            P[hd] = e / e.sum(axis=1, keepdims=True)
        mk = masks[l] if masks is not None else dict(attn=1.0, ao=1.0, f1=1.0, f2=1.0)
        Pd = P * mk["attn"]                                  # dropout on the attention weights
        for hd in range(2):
            sl = slice(25 * hd, 25 * hd + 25)
            ctx[:, sl] = Pd[hd] @ v[:, sl]
        ao = (ctx @ W[p + "self_attn.out_proj.weight"].T + W[p + "self_attn.out_proj.bias"]) * mk["ao"]
        y1, xh1, rs1 = _ln_fwd(X + ao, W[p + "norm1.weight"], W[p + "norm1.bias"])
        f1 = np.maximum(y1 @ W[p + "linear1.weight"].T + W[p + "linear1.bias"], 0) * mk["f1"]
        f2 = (f1 @ W[p + "linear2.weight"].T + W[p + "linear2.bias"]) * mk["f2"]
        y2, xh2, rs2 = _ln_fwd(y1 + f2, W[p + "norm2.weight"], W[p + "norm2.bias"])
        layers.append(dict(X=X, q=q, k=k, v=v, P=P, Pd=Pd, ctx=ctx, y1=y1, xh1=xh1, rs1=rs1, f1=f1, xh2=xh2, rs2=rs2, mk=mk))
        X = y2
    ain = np.concatenate([state[0, :6], X[0]])
    a1 = np.maximum(W["action_mlp.0.weight"] @ ain + W["action_mlp.0.bias"], 0)
    a2 = np.maximum(W["action_mlp.2.weight"] @ a1 + W["action_mlp.2.bias"], 0)
    value = float(W["action_mlp.4.weight"][0] @ a2 + W["action_mlp.4.bias"][0])
    # ---- backward ----
    dv = 2.0 * (value - float(target)) / batch
    G["action_mlp.4.weight"][0] += dv * a2
    G["action_mlp.4.bias"][0] += dv
    da2 = dv * W["action_mlp.4.weight"][0] * (a2 > 0)
    G["action_mlp.2.weight"] += np.outer(da2, a1)
    G["action_mlp.2.bias"] += da2
    da1 = (W["action_mlp.2.weight"].T @ da2) * (a1 > 0)
    G["action_mlp.0.weight"] += np.outer(da1, ain)
    G["action_mlp.0.bias"] += da1
    dX = np.zeros((T, 50))
    dX[0] = (W["action_mlp.0.weight"].T @ da1)[6:]
    for l in (2, 1, 0):
        p = "transformer_encoder.layers.%d." % l
        c = layers[l]
        ds2, dg, db = _ln_bwd(dX, c["xh2"], c["rs2"], W[p + "norm2.weight"])
        G[p + "norm2.weight"] += dg
        G[p + "norm2.bias"] += db
        mk = c["mk"]
        df2 = ds2 * mk["f2"]
        G[p + "linear2.weight"] += df2.T @ c["f1"]
        G[p + "linear2.bias"] += df2.sum(0)
        df1 = (df2 @ W[p + "linear2.weight"]) * mk["f1"] * (c["f1"] != 0)      # f1 already carries mask and ReLU
        G[p + "linear1.weight"] += df1.T @ c["y1"]
        G[p + "linear1.bias"] += df1.sum(0)
        dy1 = ds2 + df1 @ W[p + "linear1.weight"]
        ds1, dg, db = _ln_bwd(dy1, c["xh1"], c["rs1"], W[p + "norm1.weight"])
        G[p + "norm1.weight"] += dg
        G[p + "norm1.bias"] += db
        dao = ds1 * mk["ao"]
        G[p + "self_attn.out_proj.weight"] += dao.T @ c["ctx"]
        G[p + "self_attn.out_proj.bias"] += dao.sum(0)
        dctx = dao @ W[p + "self_attn.out_proj.weight"]
        dqkv = np.zeros((T, 150))
        for hd in range(2):
            sl = slice(25 * hd, 25 * hd + 25)
            Ph = c["P"][hd]
            ma = mk["attn"][hd] if np.ndim(mk["attn"]) else mk["attn"]
            dP = (dctx[:, sl] @ c["v"][:, sl].T) * ma
            dqkv[:, 100 + 25 * hd:125 + 25 * hd] = c["Pd"][hd].T @ dctx[:, sl]
            dS = Ph * (dP - (Ph * dP).sum(axis=1, keepdims=True))
            dqkv[:, sl] = dS @ c["k"][:, sl] / 5.0
            dqkv[:, 50 + 25 * hd:75 + 25 * hd] = dS.T @ c["q"][:, sl] / 5.0
        G[p + "self_attn.in_proj_weight"] += dqkv.T @ c["X"]
        G[p + "self_attn.in_proj_bias"] += dqkv.sum(0)
        dX = ds1 + dqkv @ W[p + "self_attn.in_proj_weight"]
    dA = dX / K
    dh = np.zeros((T, 50))
    cc = "in_mlp.rnn_cell."
    for st in reversed(steps):
        pk = st["p"]
        dlogit = (dA * st["hn"]).sum(1) * pk * (1 - pk)
        G["in_mlp.ponder_linear.weight"][0] += dlogit @ st["hn"]
        G["in_mlp.ponder_linear.bias"][0] += dlogit.sum()
        dh = dh + pk[:, None] * dA + dlogit[:, None] * wp[None]
        dz = dh * (st["q"] - st["h"])
        daq = dh * st["z"] * (1 - st["q"] ** 2)
        dhp = dh * (1 - st["z"])
        G[cc + "convq.weight"] += daq.T @ st["rhx"]
        G[cc + "convq.bias"] += daq.sum(0)
        drh = (daq @ W[cc + "convq.weight"])[:, :50]
        dr = drh * st["h"]
        dhp = dhp + drh * st["r"]
        for gate, dgate, a1k, a2k, gk in (("convz", dz, "az1", "az2", "z"), ("convr", dr, "ar1", "ar2", "r")):
            d2 = dgate * st[gk] * (1 - st[gk]) * (st[a2k] > 0)
            G[cc + gate + ".2.weight"] += d2.T @ st[a1k]
            G[cc + gate + ".2.bias"] += d2.sum(0)
            d1 = (d2 @ W[cc + gate + ".2.weight"]) * (st[a1k] > 0)
            G[cc + gate + ".0.weight"] += d1.T @ st["hx"]
            G[cc + gate + ".0.bias"] += d1.sum(0)
            dhp = dhp + (d1 @ W[cc + gate + ".0.weight"])[:, :50]
        dh = dhp
    return value, np.concatenate([G[k].reshape(-1) for k, _ in wts.WEIGHT_SPEC])


def dropout_masks_host(seed, ctr, b, T, p):
    """The inverted-dropout masks aem_train_grad draws for sample b (csrc/train.cu DropGen: splitmix64 finaliser of
    seed / call counter / (sample, layer, site, element)), regenerated on the host: a list of three dicts
    {"attn" [2, T, T], "ao" [T, 50], "f1" [T, 150], "f2" [T, 50]} with entries 0 or 1 / (1 - p)."""
    M = (1 << 64) - 1
    base = (int(seed) * 0x9E3779B97F4A7C15 + int(ctr) * 0xC2B2AE3D27D4EB4F) & M
    thresh = int(np.float32(p) * np.float32(16777216.0))
    scale = float(np.float32(1.0) / (np.float32(1.0) - np.float32(p)))

    def site_mask(l, site, count):
        with np.errstate(over="ignore"):
            idx = np.arange(count, dtype=np.uint64)
            key = (np.uint64(((b * 12 + l * 4 + site) << 16) & M) | idx)
            z = np.uint64(base) + key * np.uint64(0xD1B54A32D192ED03)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        keep = (z >> np.uint64(40)).astype(np.int64) >= thresh
        return np.where(keep, scale, 0.0)

    out = []
    for l in range(3):
        out.append(dict(attn=site_mask(l, 0, 2 * T * T).reshape(2, T, T), ao=site_mask(l, 1, T * 50).reshape(T, 50),
                        f1=site_mask(l, 2, T * 150).reshape(T, 150), f2=site_mask(l, 3, T * 50).reshape(T, 50)))
    return out


def batch_forward_backward(flat, states, targets, masks_of=None):
    """-> (loss, grad[113752] float64, values[B], K); masks_of(b) (optional) = the dropout masks of sample b"""
    B = len(states)
    K = batch_steps(flat, states)
    grad = np.zeros(wts.NUM_WEIGHTS)
    values = np.zeros(B)
    for b in range(B):
        values[b], g = sample_forward_backward(flat, states[b], targets[b], K, B, masks_of(b) if masks_of else None)
        grad += g
    loss = float(((values - np.asarray(targets, dtype=np.float64).reshape(-1)) ** 2).mean())
    return loss, grad, values, K
