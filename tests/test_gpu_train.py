"""GPU parity of the hand-written training step (csrc/train.cu): aem_train_grad against torch autograd of
the module mirror (float64 on the same device), aem_adam_step / the FusedStep trainer against torch.optim.Adam."""
import copy

import numpy as np
import pytest
import torch

from aemcarl_b200 import weights as wts
from test_train_math import autograd_reference, random_batch

pytestmark = pytest.mark.gpu

# float32 kernel against a float64 reference: per-tensor gradient error relative to the tensor's largest entry
GRAD_RTOL = 2e-4
VALUE_ATOL = 2e-5


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from aemcarl_b200.engine import Engine
    return Engine(0)


def run_kernel(eng, flat, states, targets):
    dev = eng.device
    p = torch.from_numpy(np.asarray(flat, dtype=np.float32)).to(dev)
    s = torch.from_numpy(np.asarray(states, dtype=np.float32)).to(dev).contiguous()
    t = torch.from_numpy(np.asarray(targets, dtype=np.float32)).to(dev)
    g = torch.full((wts.NUM_WEIGHTS,), float("nan"), device=dev)
    g, loss, values, steps = eng.train_grad(p, s, t, g)
    torch.cuda.synchronize()
    return float(loss[0]), g.cpu().numpy().astype(np.float64), values.cpu().numpy().astype(np.float64), int(steps[0])


def check_grad(grad, ref, atol=0.0, kink_frac=0.0):
    """atol: float32 noise floor for batches whose 1 / B scaling makes whole tensors' gradients ~1e-4 and smaller.
    kink_frac: a float32 kernel and a float64 reference can disagree on the side of a ReLU kink for a pre-activation within
    rounding of zero; that unit's derivative then differs (one bias element / one weight row, ~1e-3 of the tensor's largest
    entry) while everything else agrees to 2e-6.  Up to this fraction of a tensor's elements (at least one) may exceed the
    tolerance if they stay within 5e-3 of the tensor's largest entry -- an error of the algorithm moves every element."""
    assert np.isfinite(grad).all()
    for key, (off, shape) in wts.weight_offsets().items():
        n = int(np.prod(shape))
        a, b = grad[off:off + n], ref[off:off + n]
        scale = max(np.abs(b).max(), 1e-4)
        err = np.abs(a - b)
        if kink_frac > 0.0:
            out = err > GRAD_RTOL * scale + atol
            assert out.sum() <= max(1, int(kink_frac * n)) and err.max() <= 5e-3 * scale, (key, int(out.sum()), err.max(), scale)
        else:
            assert err.max() <= GRAD_RTOL * scale + atol, (key, err.max(), scale)


@pytest.mark.parametrize("ponder_bias,expect_K", [(4.0, 1), (None, 2), (-2.0, 3)])
@pytest.mark.parametrize("B,n", [(100, 5), (7, 1), (33, 10), (9, 18), (300, 5), (5, 32)])
def test_train_grad_matches_autograd(eng, ponder_bias, expect_K, B, n):
    """the trainer's batch (100 x 5 humans), one human, global-stash sizes (n = 10, 18, 32) and a batch larger than the
    grid (300: CTAs accumulate several samples into their slice), for K = 1, 2, 3 batch-global GRU steps"""
    flat = wts.synthetic_weights(3, ponder_bias=ponder_bias)
    states, targets = random_batch(B + n, B, n)
    states = states.astype(np.float32).astype(np.float64)
    targets = targets.astype(np.float32).astype(np.float64)
    loss_r, grad_r, val_r, K_r = autograd_reference(flat, states, targets, device="cuda:0")
    loss, grad, values, K = run_kernel(eng, flat, states, targets)
    assert K == K_r and (expect_K != 3 or K == 3)
    assert np.abs(values - val_r).max() < VALUE_ATOL
    assert loss == pytest.approx(loss_r, rel=1e-4)
    check_grad(grad, grad_r)


@pytest.mark.parametrize("act_steps", [1, 2, 3])
@pytest.mark.parametrize("B,n", [(100, 5), (9, 18), (300, 5)])
def test_train_grad_act_fixed_matches_autograd(eng, act_steps, B, n):
    """act_fixed (components.py:109-112): exactly act_steps GRU steps, h_K handed on, nothing flows into ponder_linear"""
    from aemcarl_b200.crowd_nav.common.value_network import ValueNetworkTransformerAndGRU
    flat = wts.synthetic_weights(4, ponder_bias=None)
    states, targets = random_batch(1000 + B + n + act_steps, B, n)
    states = states.astype(np.float32).astype(np.float64)
    targets = targets.astype(np.float32).astype(np.float64)
    model = ValueNetworkTransformerAndGRU(act_fixed=True, act_steps=act_steps).double().eval()
    model.load_flat_weights(flat)
    model = model.to("cuda:0")
    out, _ = model.forward_torch(torch.as_tensor(states, dtype=torch.float64, device="cuda:0"))
    loss_t = torch.nn.functional.mse_loss(out, torch.as_tensor(targets, dtype=torch.float64, device="cuda:0").view(-1, 1))
    loss_t.backward()
    named = dict(model.named_parameters())
    grad_r = np.concatenate([(named[k].grad if named[k].grad is not None else torch.zeros_like(named[k])).detach().cpu().numpy()
                             .reshape(-1) for k, _ in wts.WEIGHT_SPEC])
    try:
        eng.set_act_mode(True, act_steps)
        loss, grad, values, K = run_kernel(eng, flat, states, targets)
    finally:
        eng.set_act_mode(False, 1)
    assert K == act_steps
    assert np.abs(values - out.detach().cpu().numpy().reshape(-1)).max() < VALUE_ATOL
    assert loss == pytest.approx(float(loss_t), rel=1e-4)
    check_grad(grad, grad_r, atol=1e-7, kink_frac=0.01)     # (atol: the r gate's gradients are below 1e-4 here)
    off, shape = wts.weight_offsets()["in_mlp.ponder_linear.weight"] if "in_mlp.ponder_linear.weight" in wts.weight_offsets() \
        else next(v for k, v in wts.weight_offsets().items() if "ponder" in k and "weight" in k)
    assert (grad[off:off + int(np.prod(shape))] == 0).all()


@pytest.mark.parametrize("p,B,n,ponder_bias", [(0.1, 24, 5, None), (0.1, 9, 10, -2.0), (0.3, 16, 5, 4.0), (0.1, 200, 5, None)])
def test_train_grad_with_live_dropout_matches_the_masked_derivation(eng, p, B, n, ponder_bias):
    """The reference trains with nn.TransformerEncoderLayer's dropout live (qnetwork_factory.py:628, trainer.py:65-86).
    aem_set_train_dropout(p, seed) switches the four sites on; the kernel's counter-based masks are regenerated on the host
    (manual_backward.dropout_masks_host) and injected into the float64 derivation, which equals torch autograd of the masked
    forward (tests/test_train_math.py): same gradients, values and loss; a second call draws fresh masks; p = 0 is eval()."""
    import manual_backward as mb
    flat = wts.synthetic_weights(6, ponder_bias=ponder_bias)
    states, targets = random_batch(40 + B + n, B, n)
    states = states.astype(np.float32).astype(np.float64)
    targets = targets.astype(np.float32).astype(np.float64)
    seed = 1234567
    T = n + 1
    try:
        eng.set_train_dropout(p, seed)
        out = [run_kernel(eng, flat, states, targets) for _ in range(2)]          # call counters 0 and 1
        for ctr, (loss, grad, values, K) in enumerate(out):
            loss_r, grad_r, val_r, K_r = mb.batch_forward_backward(
                flat, states, targets, masks_of=lambda b: mb.dropout_masks_host(seed, ctr, b, T, p))
            assert K == K_r
            assert np.abs(values - val_r).max() < VALUE_ATOL
            assert loss == pytest.approx(loss_r, rel=1e-4)
            check_grad(grad, grad_r, atol=1e-7)
        assert not np.array_equal(out[0][1], out[1][1])                           # fresh masks per step
        keep = np.concatenate([m["f1"].ravel() for b in range(B) for m in mb.dropout_masks_host(seed, 0, b, T, p)])
        assert abs((keep == 0).mean() - p) < 0.02                                 # the drop rate is p
    finally:
        eng.set_train_dropout(0.0, 0)
    loss0, grad0, values0, K0 = run_kernel(eng, flat, states, targets)
    loss_r, grad_r, val_r, K_r = mb.batch_forward_backward(flat, states, targets)
    check_grad(grad0, grad_r, atol=1e-7)


def test_fused_trainer_follows_the_module_mode(eng):
    """Trainer(fused=True) in train() mode optimises with dropout (the reference's regime), in eval() mode without"""
    from aemcarl_b200.crowd_nav.default_configs import policy_config
    from aemcarl_b200.crowd_nav.policy.policy_factory import policy_factory
    from aemcarl_b200.crowd_nav.utils.trainer import Trainer
    dev = torch.device("cuda:0")
    policy = policy_factory["actenvcarl"]()
    policy.configure(policy_config())
    policy.set_device(dev)
    model = policy.get_model()
    x = torch.randn((100, 5, 13), device=dev)
    v = torch.randn((100, 1), device=dev)
    tr = Trainer(model, None, dev, 100, fused=True)
    tr.set_learning_rate(1e-3, "Adam")
    model.eval()
    l_eval = [float(tr._step(x, v)) for _ in range(3)]
    assert tr._fused._drop_p == 0.0
    model.train()
    l_train = [float(tr._step(x, v)) for _ in range(3)]
    assert tr._fused._drop_p == pytest.approx(0.1)
    model.eval()
    float(tr._step(x, v))
    assert tr._fused._drop_p == 0.0
    assert all(np.isfinite(l_eval + l_train))


def test_train_grad_is_deterministic(eng):
    flat = wts.synthetic_weights(5)
    states, targets = random_batch(2, 100, 5)
    a = run_kernel(eng, flat, states, targets)
    b = run_kernel(eng, flat, states, targets)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


def test_mixed_halting_batch_uses_the_batch_global_count(eng):
    """components.py:132-134: one sample that needs a third step makes the whole batch take it"""
    flat = wts.synthetic_weights(4, ponder_bias=-0.13)
    states, targets = random_batch(11, 64, 5)
    loss_r, grad_r, val_r, K_r = autograd_reference(flat, states, targets, device="cuda:0")
    loss, grad, values, K = run_kernel(eng, flat, states, targets)
    assert K == K_r
    assert np.abs(values - val_r).max() < VALUE_ATOL
    check_grad(grad, grad_r)


def test_adam_step_equals_torch_adam(eng):
    dev = eng.device
    g = torch.Generator(device=dev).manual_seed(1)
    p0 = torch.randn(wts.NUM_WEIGHTS, device=dev, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.99), eps=1e-3)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for step in range(1, 6):
        grad = torch.randn(wts.NUM_WEIGHTS, device=dev, generator=g) * 0.1
        ref.grad = grad.clone()
        opt.step()
        eng.adam_step(p, grad, m, v, 1e-3, (0.9, 0.99), 1e-3, step)
        assert torch.allclose(p, ref.detach(), rtol=2e-7, atol=2e-7), float((p - ref.detach()).abs().max())   # 1 ulp
    assert torch.allclose(m, opt.state[ref]["exp_avg"], rtol=1e-5, atol=1e-8)
    assert torch.allclose(v, opt.state[ref]["exp_avg_sq"], rtol=1e-5, atol=1e-10)
    # a summed gradient of two ranks, averaged by grad_scale
    p2, m2, v2 = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    p3, m3, v3 = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    grad = torch.randn(wts.NUM_WEIGHTS, device=dev, generator=g)
    eng.adam_step(p2, grad * 2, m2, v2, 1e-3, (0.9, 0.99), 1e-3, 1, grad_scale=0.5)
    eng.adam_step(p3, grad, m3, v3, 1e-3, (0.9, 0.99), 1e-3, 1)
    assert torch.equal(p2, p3)


def test_fused_trainer_walks_with_the_eager_trainer():
    """Trainer(fused=True) against the eager torch step on the same six mini-batches: same losses, same parameters; the
    module's parameters are views into the flat vector, the lookahead engine sees the updated weights, and the torch
    optimizer's state describes the fused moments."""
    from aemcarl_b200.crowd_nav.utils.trainer import Trainer
    from test_gpu_api import setup
    env, robot, policy = setup(seed=4, ponder_bias=-0.13)
    dev = torch.device("cuda:0")
    model_a = policy.get_model()
    model_b = copy.deepcopy(model_a).to(dev)
    g = torch.Generator(device=dev).manual_seed(5)
    batches = [(torch.randn((100, 5, 13), device=dev, generator=g), torch.randn((100, 1), device=dev, generator=g))
               for _ in range(6)]
    losses = {}
    trainers = {}
    for name, model, fused in (("eager", model_a, False), ("fused", model_b, True)):
        tr = Trainer(model, None, dev, 100, fused=fused)
        tr.set_learning_rate(1e-3, "Adam")
        losses[name] = [float(tr._step(x, v)) for x, v in batches]
        trainers[name] = tr
    assert np.allclose(losses["eager"], losses["fused"], rtol=2e-4, atol=1e-6), (losses["eager"], losses["fused"])
    for (k, pa), pb in zip(model_a.named_parameters(), model_b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-3, atol=2e-5), (k, float((pa - pb).abs().max()))
    assert losses["fused"][0] != losses["fused"][-1]
    fs = trainers["fused"]._fused
    assert fs.step == 6
    flat_now = torch.from_numpy(model_b.flat_weights()).to(dev)
    assert torch.equal(flat_now, fs.flat)                       # parameters ARE the flat vector
    fs.export_state()
    st = trainers["fused"].optimizer.state[next(model_b.parameters())]
    assert float(st["step"]) == 6 and st["exp_avg"].data_ptr() == fs.exp_avg.data_ptr()
    # the value kernel of the policy path picks the new weights up (version bump through note_external_update)
    x = batches[0][0][:1]
    with torch.no_grad():
        v_kernel = float(model_b(x)[0])
        v_torch = float(model_b.forward_torch(x)[0])
    assert v_kernel == pytest.approx(v_torch, abs=2e-5)
    # one eager step of the same optimizer continues from the fused state
    tr = trainers["fused"]
    tr.fused = False
    before = fs.flat.clone()
    tr._step(*batches[0])
    assert not torch.equal(before, torch.from_numpy(model_b.flat_weights()).to(dev))


def test_sample_batch_draws_a_uniform_subset_without_replacement(eng):
    """aem_sample_batch against its contract: k distinct indices below `size`, the gathered rows, reproducible per (seed, counter),
    different across counters, and uniform over the ring (chi-square over 200 bins of 2,000 batches)."""
    dev = eng.device
    cap, size, k, n = 5000, 4000, 100, 5
    states = torch.arange(cap, device=dev, dtype=torch.float32).view(cap, 1, 1).expand(cap, n, 13).contiguous()
    values = -torch.arange(cap, device=dev, dtype=torch.float32).view(cap, 1)
    s, v, idx = eng.sample_batch(states, values, size, k, 7, 1, want_index=True)
    assert idx.min() >= 0 and idx.max() < size and len(torch.unique(idx)) == k
    assert torch.equal(s, states[idx]) and torch.equal(v, values[idx])
    s2, v2, idx2 = eng.sample_batch(states, values, size, k, 7, 1, want_index=True)
    assert torch.equal(idx, idx2)
    assert not torch.equal(idx, eng.sample_batch(states, values, size, k, 7, 2, want_index=True)[2])
    counts = torch.zeros(size, device=dev)
    for c in range(2000):
        counts[eng.sample_batch(states, values, size, k, 11, c, want_index=True)[2]] += 1
    bins = counts.view(200, -1).sum(1).cpu().numpy()
    expect = 2000 * k / 200
    chi2 = float(((bins - expect) ** 2 / expect).sum())
    assert chi2 < 300, chi2                                  # 199 degrees of freedom: mean 199, sd 20
    full = eng.sample_batch(states, values, 64, 64, 3, 9, want_index=True)[2]     # k == size: a permutation of everything
    assert sorted(full.tolist()) == list(range(64))


def test_sgd_and_rmsprop_steps_equal_torch(eng):
    """aem_sgd_step / aem_rmsprop_step against torch.optim on the same gradient sequence (the other two optimizers of
    Trainer.set_learning_rate, trainer.py:24-34)"""
    dev = eng.device
    g = torch.Generator(device=dev).manual_seed(2)
    p0 = torch.randn(wts.NUM_WEIGHTS, device=dev, generator=g)
    for kind in ("sgd", "rmsprop"):
        ref = torch.nn.Parameter(p0.clone())
        opt = torch.optim.SGD([ref], lr=1e-3, momentum=0.9) if kind == "sgd" else torch.optim.RMSprop([ref], lr=1e-3, alpha=0.9)
        p, buf = p0.clone(), torch.zeros_like(p0)
        for step in range(1, 6):
            grad = torch.randn(wts.NUM_WEIGHTS, device=dev, generator=g) * 0.1
            ref.grad = grad.clone()
            opt.step()
            if kind == "sgd":
                eng.sgd_step(p, grad, buf, 1e-3, 0.9, step)
            else:
                eng.rmsprop_step(p, grad, buf, 1e-3, 0.9, 1e-8)
            assert torch.allclose(p, ref.detach(), rtol=1e-6, atol=1e-6), (kind, step, float((p - ref.detach()).abs().max()))
        state = opt.state[ref]["momentum_buffer" if kind == "sgd" else "square_avg"]
        assert torch.allclose(buf, state, rtol=1e-5, atol=1e-7), kind


def test_fused_sgd_walks_with_torch():
    """Trainer(fused=True) with optim.SGD(momentum 0.9) against the eager step on the same five mini-batches.  (RMSprop divides by
    sqrt(g^2) on its first step: rounding-level gradient differences become sign flips of +-lr, so only its update rule is
    compared, above.)"""
    optimizer = "SGD"
    from aemcarl_b200.crowd_nav.utils.trainer import Trainer
    from test_gpu_api import setup
    env, robot, policy = setup(seed=4, ponder_bias=-0.13)
    dev = torch.device("cuda:0")
    model_a = policy.get_model()
    model_b = copy.deepcopy(model_a).to(dev)
    g = torch.Generator(device=dev).manual_seed(6)
    batches = [(torch.randn((100, 5, 13), device=dev, generator=g), torch.randn((100, 1), device=dev, generator=g))
               for _ in range(5)]
    losses = {}
    for name, model, fused in (("eager", model_a, False), ("fused", model_b, True)):
        tr = Trainer(model, None, dev, 100, fused=fused)
        tr.set_learning_rate(1e-3, optimizer)
        losses[name] = [float(tr._step(x, v)) for x, v in batches]
    assert np.allclose(losses["eager"], losses["fused"], rtol=2e-4, atol=1e-6), (losses["eager"], losses["fused"])
    for (k, pa), pb in zip(model_a.named_parameters(), model_b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-3, atol=2e-5), (k, float((pa - pb).abs().max()))
    assert losses["fused"][0] != losses["fused"][-1]
    # RMSprop runs through the same trainer path
    tr = Trainer(copy.deepcopy(model_a).to(dev), None, dev, 100, fused=True)
    tr.set_learning_rate(1e-3, "RMSprop")
    assert all(np.isfinite(float(tr._step(x, v))) for x, v in batches)


def test_optimize_batch_single_call_and_graph_replay():
    """Trainer.optimize_batch on a DeviceReplayMemory with the fused Adam step: (a) one C call per iteration with the step state
    on the device (aem_train_step_adam) walks like sample() + step; (b) replaying the CUDA graph of that call gives bit-identical
    parameters; (c) a changed ring size re-captures."""
    from aemcarl_b200.crowd_nav.utils.trainer import Trainer
    from aemcarl_b200.crowd_nav.utils.vec_explorer import DeviceReplayMemory
    from test_gpu_api import setup
    env, robot, policy = setup(seed=4, ponder_bias=-0.13)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(9)
    states = torch.randn((5000, 5, 13), device=dev, generator=g)
    values = torch.randn((5000, 1), device=dev, generator=g)
    flats = {}
    for name, graph, single in (("stepwise", False, False), ("single_call", False, True), ("graph", True, True)):
        mem = DeviceReplayMemory(8000, 5, dev)
        mem.push_batch(states[:4000], values[:4000])
        model = copy.deepcopy(policy.get_model()).to(dev)
        tr = Trainer(model, mem, dev, 100, fused=True, cuda_graph=graph)
        tr.set_learning_rate(1e-3, "Adam")
        if single:
            l1 = tr.optimize_batch(12)
            mem.push_batch(states[4000:], values[4000:])            # the ring grows: the graph is re-captured
            l2 = tr.optimize_batch(8)
        else:
            losses = [float(tr._step(*mem.sample(100))) for _ in range(12)]
            l1 = float(np.mean(losses))
            mem.push_batch(states[4000:], values[4000:])
            l2 = float(np.mean([float(tr._step(*mem.sample(100))) for _ in range(8)]))
        assert np.isfinite(l1) and np.isfinite(l2)
        assert tr._fused.step == 20 and mem._samples_drawn == 20
        flats[name] = (tr._fused.flat.clone(), l1, l2)
    assert torch.equal(flats["single_call"][0], flats["graph"][0])
    assert flats["single_call"][1:] == flats["graph"][1:]
    assert torch.allclose(flats["stepwise"][0], flats["single_call"][0], rtol=1e-5, atol=1e-6)
    assert flats["stepwise"][1] == pytest.approx(flats["single_call"][1], rel=1e-5)


def test_fused_trainer_with_act_fixed_walks_with_the_eager_trainer():
    """act_fixed = True, act_steps = 2 (train.py --act_fixed --act_steps 2): the fused step against the eager torch step"""
    from aemcarl_b200.crowd_nav.common.value_network import ValueNetworkTransformerAndGRU
    from aemcarl_b200.crowd_nav.utils.trainer import Trainer
    dev = torch.device("cuda:0")
    model_a = ValueNetworkTransformerAndGRU(act_fixed=True, act_steps=2).eval()
    model_a.load_flat_weights(wts.synthetic_weights(8))
    model_a = model_a.to(dev)
    model_b = copy.deepcopy(model_a)
    g = torch.Generator(device=dev).manual_seed(9)
    batches = [(torch.randn((100, 5, 13), device=dev, generator=g), torch.randn((100, 1), device=dev, generator=g))
               for _ in range(5)]
    losses = {}
    for name, model, fused in (("eager", model_a, False), ("fused", model_b, True)):
        tr = Trainer(model, None, dev, 100, fused=fused)
        tr.set_learning_rate(1e-3, "Adam")
        losses[name] = [float(tr._step(x, v)) for x, v in batches]
    assert np.allclose(losses["eager"], losses["fused"], rtol=2e-4, atol=1e-6), (losses["eager"], losses["fused"])
    for (k, pa), pb in zip(model_a.named_parameters(), model_b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-3, atol=2e-5), (k, float((pa - pb).abs().max()))
    assert losses["fused"][0] != losses["fused"][-1]
