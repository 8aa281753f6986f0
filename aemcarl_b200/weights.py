"""Flat weight blob of the flag-5 value network (ValueNetworkTransformerAndGRU).

The C ABI (include/aemcarl_b200.h, aem_load_weights) takes the LIVE parameters of the reference module
(crowd_nav/common/qnetwork_factory.py:586-639) as one float32 array in state_dict order, without the dead
`encoder_layer.*` template: 113,752 floats.
"""
import numpy as np

NUM_WEIGHTS = 113752

# (state_dict key, shape) in blob order
WEIGHT_SPEC = []
for _g in ("convz", "convr"):
    WEIGHT_SPEC += [("in_mlp.rnn_cell.%s.0.weight" % _g, (100, 63)), ("in_mlp.rnn_cell.%s.0.bias" % _g, (100,)),
                    ("in_mlp.rnn_cell.%s.2.weight" % _g, (50, 100)), ("in_mlp.rnn_cell.%s.2.bias" % _g, (50,))]
WEIGHT_SPEC += [("in_mlp.rnn_cell.convq.weight", (50, 63)), ("in_mlp.rnn_cell.convq.bias", (50,)),
                ("in_mlp.ponder_linear.weight", (1, 50)), ("in_mlp.ponder_linear.bias", (1,))]
for _l in range(3):
    _p = "transformer_encoder.layers.%d." % _l
    WEIGHT_SPEC += [(_p + "self_attn.in_proj_weight", (150, 50)), (_p + "self_attn.in_proj_bias", (150,)),
                    (_p + "self_attn.out_proj.weight", (50, 50)), (_p + "self_attn.out_proj.bias", (50,)),
                    (_p + "linear1.weight", (150, 50)), (_p + "linear1.bias", (150,)),
                    (_p + "linear2.weight", (50, 150)), (_p + "linear2.bias", (50,)),
                    (_p + "norm1.weight", (50,)), (_p + "norm1.bias", (50,)),
                    (_p + "norm2.weight", (50,)), (_p + "norm2.bias", (50,))]
WEIGHT_SPEC += [("action_mlp.0.weight", (100, 56)), ("action_mlp.0.bias", (100,)),
                ("action_mlp.2.weight", (50, 100)), ("action_mlp.2.bias", (50,)),
                ("action_mlp.4.weight", (1, 50)), ("action_mlp.4.bias", (1,))]
assert sum(int(np.prod(s)) for _, s in WEIGHT_SPEC) == NUM_WEIGHTS

# dead template keys present in reference checkpoints (kept for load_state_dict compatibility)
DEAD_TEMPLATE_SPEC = [("encoder_layer." + k.split("layers.0.")[1], s) for k, s in WEIGHT_SPEC
                      if k.startswith("transformer_encoder.layers.0.")]


def weight_offsets():
    off, out = 0, {}
    for k, s in WEIGHT_SPEC:
        out[k] = (off, s)
        off += int(np.prod(s))
    return out


def flatten_state_dict(sd):
    """state_dict (torch tensors or arrays) -> float32[113752] in blob order."""
    parts = []
    for k, s in WEIGHT_SPEC:
        v = sd[k]
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        v = np.asarray(v, dtype=np.float32)
        assert tuple(v.shape) == tuple(s), (k, v.shape, s)
        parts.append(v.reshape(-1))
    return np.concatenate(parts)


def unflatten(flat):
    """float32[113752] -> {key: ndarray} (views)."""
    flat = np.asarray(flat, dtype=np.float32)
    assert flat.size == NUM_WEIGHTS
    return {k: flat[o:o + int(np.prod(s))].reshape(s) for k, (o, s) in weight_offsets().items()}


def synthetic_weights(seed=0, ponder_bias=None, gain=1.0):
    """Deterministic random-init weights from numpy's legacy MT19937 stream (stable across numpy/torch
    versions, so golden fixtures need not store 455 KB per weight set).  nn.Linear-style U(-1/sqrt(fan_in), ..);
    LayerNorm gamma = 1 + 0.1 u, beta = 0.1 u.  `ponder_bias` overrides in_mlp.ponder_linear.bias to steer the
    adaptive-computation-time halting count (components.py:125-134)."""
    rng = np.random.RandomState(seed)
    parts = []
    for k, s in WEIGHT_SPEC:
        if "norm" in k:
            u = rng.uniform(-1.0, 1.0, size=s)
            v = (1.0 + 0.1 * u) if k.endswith("weight") else 0.1 * u
        else:
            fan_in = s[1] if len(s) == 2 else {100: 63, 50: 100, 150: 50, 1: 50}.get(s[0], 50)
            b = gain / np.sqrt(fan_in)
            v = rng.uniform(-b, b, size=s)
        if k == "in_mlp.ponder_linear.bias" and ponder_bias is not None:
            v = np.full(s, ponder_bias)
        parts.append(np.asarray(v, dtype=np.float32).reshape(-1))
    return np.concatenate(parts)
