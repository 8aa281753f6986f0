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
