"""CPU check of the hand derivation csrc/train.cu implements (tests/manual_backward.py) against torch autograd of the
module mirror (float64), for every batch-global ACT step count."""
import numpy as np
import pytest
import torch

from aemcarl_b200 import weights as wts
from aemcarl_b200.crowd_nav.common.value_network import ValueNetworkTransformerAndGRU
import manual_backward as mb


def autograd_reference(flat, states, targets, dtype=torch.float64, device="cpu"):
    model = ValueNetworkTransformerAndGRU().to(dtype).eval()
    model.load_flat_weights(flat)
    model = model.to(device)
    out, _ = model.forward_torch(torch.as_tensor(states, dtype=dtype, device=device))
    loss = torch.nn.functional.mse_loss(out, torch.as_tensor(targets, dtype=dtype, device=device).view(-1, 1))
    loss.backward()
    named = dict(model.named_parameters())
    grad = np.concatenate([named[k].grad.detach().cpu().numpy().reshape(-1) for k, _ in wts.WEIGHT_SPEC])
    return float(loss), grad, out.detach().cpu().numpy().reshape(-1), int(model.step_cnt)


def random_batch(seed, B, n):
    rng = np.random.RandomState(seed)
    states = rng.uniform(-1, 1, size=(B, n, 13))
    states[:, :, :6] = states[:, :1, :6]            # the robot part repeats in every row of a rotated state
    return states, rng.uniform(-1, 1, size=(B,))


@pytest.mark.parametrize("ponder_bias,expect_K", [(4.0, 1), (None, 2), (-2.0, 3)])
def test_manual_backward_equals_autograd(ponder_bias, expect_K):
    flat = wts.synthetic_weights(3, ponder_bias=ponder_bias)
    states, targets = random_batch(1, 7, 5)
    loss_r, grad_r, val_r, K_r = autograd_reference(flat, states, targets)
    loss_m, grad_m, val_m, K_m = mb.batch_forward_backward(flat, states, targets)
    assert K_m == K_r == expect_K
    assert loss_m == pytest.approx(loss_r, rel=1e-12)
    assert np.allclose(val_m, val_r, atol=1e-12)
    assert np.abs(grad_m - grad_r).max() < 1e-12 * max(1.0, np.abs(grad_r).max())
    assert np.abs(grad_r).max() > 1e-3


def _torch_forward_with_masks(model, states, masks):
    """the module's forward with EXPLICIT inverted-dropout masks at the four sites of every nn.TransformerEncoderLayer
    (attention weights, attention branch, after the FFN ReLU, FFN branch) -- what F.dropout does in train() mode, with the
    random draw replaced by given masks"""
    B, n, _ = states.shape
    T = n + 1
    x = torch.cat([model._robot_token(states), states], dim=1)
    h0 = torch.zeros([B * T, 50], dtype=states.dtype)
    out, K = model.in_mlp(x, h0)
    X = out.view(B, T, 50)
    for l, layer in enumerate(model.transformer_encoder.layers):
        mk = {k: torch.as_tensor(np.stack([masks[b][l][k] for b in range(B)])) for k in ("attn", "ao", "f1", "f2")}
        qkv = X @ layer.self_attn.in_proj_weight.T + layer.self_attn.in_proj_bias
        q, k, v = qkv[..., :50], qkv[..., 50:100], qkv[..., 100:]
        ctx = []
        for hd in range(2):
            sl = slice(25 * hd, 25 * hd + 25)
            P = torch.softmax(q[..., sl] @ k[..., sl].transpose(1, 2) / 5.0, dim=-1) * mk["attn"][:, hd]
            ctx.append(P @ v[..., sl])
        ao = (torch.cat(ctx, dim=-1) @ layer.self_attn.out_proj.weight.T + layer.self_attn.out_proj.bias) * mk["ao"]
        y1 = layer.norm1(X + ao)
        f1 = torch.relu(layer.linear1(y1)) * mk["f1"]
        X = layer.norm2(y1 + layer.linear2(f1) * mk["f2"])
    value = model.action_mlp(torch.cat([states[:, 0, :6], X[:, 0, :]], dim=1))
    return value, K


def test_manual_backward_with_dropout_masks_equals_autograd():
    """groundwork for a dropout-active fused training step (the reference trains with dropout live): the hand derivation with the
    four masks per encoder layer against autograd of the same masked forward; with all-ones masks it is the eval() network"""
    flat = wts.synthetic_weights(3)
    B, n, p = 5, 5, 0.1
    T = n + 1
    states, targets = random_batch(4, B, n)
    rng = np.random.RandomState(0)
    draw = lambda shape: (rng.uniform(size=shape) >= p) / (1 - p)
    masks = [[dict(attn=draw((2, T, T)), ao=draw((T, 50)), f1=draw((T, 150)), f2=draw((T, 50))) for _ in range(3)] for _ in range(B)]
    model = ValueNetworkTransformerAndGRU().double().eval()
    model.load_flat_weights(flat)
    value, K = _torch_forward_with_masks(model, torch.as_tensor(states), masks)
    loss = torch.nn.functional.mse_loss(value, torch.as_tensor(targets).view(-1, 1))
    loss.backward()
    named = dict(model.named_parameters())
    grad_r = np.concatenate([named[k].grad.numpy().reshape(-1) for k, _ in wts.WEIGHT_SPEC])
    grad_m = np.zeros(wts.NUM_WEIGHTS)
    values = np.zeros(B)
    for b in range(B):
        values[b], g = mb.sample_forward_backward(flat, states[b], targets[b], int(K), B, masks=masks[b])
        grad_m += g
    assert np.allclose(values, value.detach().numpy().reshape(-1), atol=1e-12)
    assert np.abs(grad_m - grad_r).max() < 1e-12 * max(1.0, np.abs(grad_r).max())
    # all-ones masks: the eval()-mode network the fused step implements today
    ones = [[dict(attn=np.ones((2, T, T)), ao=np.ones((T, 50)), f1=np.ones((T, 150)), f2=np.ones((T, 50))) for _ in range(3)]
            for _ in range(B)]
    v1, _ = mb.sample_forward_backward(flat, states[0], targets[0], int(K), B, masks=ones[0])
    v0, _ = mb.sample_forward_backward(flat, states[0], targets[0], int(K), B)
    assert v1 == v0
