"""GPU: the three-launch decision step.  aem_env_prepare must equal aem_orca + aem_env_lookahead and aem_env_finish must
equal select + aem_env_apply + aem_env_reset_done BIT FOR BIT (they are the same arithmetic in one launch each), so every
parity statement the separate entry points carry against the oracle transfers; the oracle is checked directly as well."""
import numpy as np
import pytest
import torch

from aemcarl_b200 import scenarios, weights as wts
from aemcarl_b200.engine import Engine, make_env_params
from aemcarl_b200.vec_env import VecCrowdSim

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def scene(rule, n, B, seed, mixed=False):
    rng = np.random.RandomState(seed)
    if mixed:
        robot, humans, counts = scenarios.generate_mixed_pool("train", 10, B)
    else:
        robot, humans = scenarios.generate_pool("train", 10, B, n, rule)
        counts = None
    humans[:, :, 2:4] = rng.uniform(-0.8, 0.8, humans[:, :, 2:4].shape)        # mid-episode: everybody is moving
    robot[:, 0:2] += rng.uniform(-1, 3, (B, 2))
    robot[:, 2:4] = rng.uniform(-0.5, 0.5, (B, 2))
    if counts is not None:
        for b in range(B):
            humans[b, counts[b]:, 2:4] = 0.0
    gt = rng.uniform(0, 24.9, B)
    return robot, humans, gt, counts


@pytest.mark.parametrize("rule,n,visible,mixed", [("circle_crossing", 5, False, False), ("circle_crossing", 5, True, False),
                                                  ("square_crossing", 10, False, False), ("circle_crossing", 18, True, False),
                                                  ("circle_crossing", 1, False, False), ("mixed", 5, False, True)])
def test_env_prepare_equals_orca_then_env_lookahead(oracle, actions, rule, n, visible, mixed):
    B = 257
    robot, humans, gt, counts = scene(rule, n, B, 3, mixed)
    eng = Engine(0)
    eng.set_action_space(actions)
    eng.set_env_params(make_env_params(robot_visible=visible))
    eng.set_active_counts(dev(counts) if mixed else None, None)
    r, h, t = dev(robot), dev(humans), dev(gt)
    nv = eng.orca(h, r)
    rew, info, dmin, hnext = eng.env_lookahead(r, h, nv, t)
    nv2, rew2, info2, dmin2, hnext2 = eng.env_prepare(r, h, t)
    torch.cuda.synchronize()
    for a, b, name in ((nv, nv2, "new_vel"), (rew, rew2, "rewards"), (info, info2, "info"), (dmin, dmin2, "dmin"),
                       (hnext, hnext2, "humans_next")):
        assert np.array_equal(a.cpu().numpy(), b.cpu().numpy(), equal_nan=True), name
    if not mixed:                                            # and directly against the oracle
        p = oracle.default_params(robot_visible=int(visible))
        onv = oracle.orca(humans[:64], robot[:64], p)
        assert np.array_equal(nv2.cpu().numpy()[:64], onv)
        orew, oinfo, odmin, ohnext = oracle.env_lookahead(robot[:64], humans[:64], onv, gt[:64], actions, p)
        assert np.array_equal(info2.cpu().numpy()[:64], oinfo) and np.array_equal(hnext2.cpu().numpy()[:64], ohnext)
        assert np.abs(rew2.cpu().numpy()[:64] - orew).max() <= 1e-12


@pytest.mark.parametrize("with_pool", [False, True])
@pytest.mark.parametrize("mixed", [False, True])
def test_env_finish_equals_select_apply_reset(actions, with_pool, mixed):
    """values with exact ties, NaNs, all-NaN rows and robots already at their goal; a third of the environments end"""
    B, n, A = 333, 5, actions.shape[0]
    rng = np.random.RandomState(11)
    robot, humans, gt, counts = scene("mixed" if mixed else "circle_crossing", n, B, 5, mixed)
    robot[::17, 0:2] = robot[::17, 5:7]                       # reach_destination -> action 0
    values = np.round(rng.uniform(-1, 1, (B, A)), 1)          # one decimal: many exact ties
    values[rng.rand(B, A) < 0.05] = np.nan
    values[5] = np.nan
    info = rng.choice([0, 1, 2, 3, 4], size=(B, A), p=[0.5, 0.2, 0.1, 0.1, 0.1]).astype(np.int32)
    rewards = rng.uniform(-0.3, 1, (B, A))
    dmin = rng.uniform(0, 2, (B, A))
    new_vel = rng.uniform(-1, 1, (B, n, 2))
    pool_robot, pool_humans = scenarios.generate_pool("val", 0, 50, n, "circle_crossing")
    pool_counts = rng.randint(1, n + 1, 50).astype(np.int32)
    eng = Engine(0)
    eng.set_action_space(actions)
    eng.set_env_params(make_env_params())

    def run(fused):
        r, h, t = dev(robot), dev(humans), dev(gt)
        nact = dev(counts) if mixed else None
        eng.set_active_counts(nact, dev(pool_counts) if mixed else None)
        case_idx = dev((np.arange(B) * 3 - 7).astype(np.int32))
        outcomes = torch.zeros(5, dtype=torch.int64, device="cuda")
        v, rw, inf, dm, nv = dev(values), dev(rewards), dev(info), dev(dmin), dev(new_vel)
        if fused:
            best = torch.empty(B, dtype=torch.int32, device="cuda")
            pool = (dev(pool_robot), dev(pool_humans), case_idx, 4, outcomes) if with_pool else None
            out = eng.env_finish(v, r, h, nv, t, rw, inf, dm, best=best, pool=pool)
        else:
            best = torch.empty(B, dtype=torch.int32, device="cuda")
            # the argmax alone, on the host with the documented rule (first strict maximum, NaN never wins, goal -> 0):
            # aem_lookahead's own select is compared with env_finish in the episode test below
            hv = values
            hb = np.full(B, -1, np.int32)
            for b in range(B):
                bv = -np.inf
                for a in range(A):
                    if hv[b, a] > bv:
                        bv, hb[b] = hv[b, a], a
                d = robot[b, 0:2] - robot[b, 5:7]
                if np.sqrt(d[1] * d[1] + d[0] * d[0]) < robot[b, 4]:
                    hb[b] = 0
            best = dev(hb)
            out = eng.env_apply(r, h, nv, t, best, rw, inf, dm)
            if with_pool:
                eng.env_reset_done(r, h, t, out[1], out[2], dev(pool_robot), dev(pool_humans), case_idx, 4, outcomes)
        torch.cuda.synchronize()
        res = [x.cpu().numpy() for x in (best, r, h, t, case_idx, outcomes) + tuple(out)]
        if mixed:
            res.append(nact.cpu().numpy())
        return res

    a, b = run(False), run(True)
    names = ["best", "robot", "humans", "time", "case_idx", "outcomes", "reward", "done", "info", "dmin", "n_active"]
    for x, y, name in zip(a, b, names):
        assert np.array_equal(x, y, equal_nan=True), name
    assert a[7].sum() > B // 5 and (a[0] == -1).any() and (a[0][::17] == 0).all()


@pytest.mark.parametrize("rule,n,mixed", [("circle_crossing", 5, False), ("square_crossing", 10, False), ("mixed", 5, True)])
def test_three_launch_step_equals_the_separate_calls_over_an_episode(actions, rule, n, mixed):
    """VecCrowdSim.decide_and_step (3 launches) against lookahead() + apply() (orca/env fused + select + apply + reset) for 60
    steps with re-seeding: identical state, outcomes and launch count 3 per step"""
    B = 128
    W = wts.synthetic_weights(3, -0.05)

    def make():
        eng = Engine(0)
        eng.set_action_space(actions)
        eng.set_env_params(make_env_params())
        eng.load_weights(W)
        eng.set_precision("fp16x3")
        return eng, VecCrowdSim(eng, B, n, rule=rule, phase="test", pool_size=300)

    e1, v1 = make()
    e2, v2 = make()
    for step in range(60):
        v1.lookahead()
        o1 = v1.apply()
        c0 = e2.launch_count()
        o2 = v2.decide_and_step()
        assert e2.launch_count() - c0 == 3
        for x, y in zip(o1, o2):
            assert torch.equal(x, y), step
        assert torch.equal(v1.best, v2.best) and torch.equal(v1.robot, v2.robot) and torch.equal(v1.humans, v2.humans)
        assert torch.equal(v1.global_time, v2.global_time) and torch.equal(v1.case_idx, v2.case_idx)
    assert torch.equal(v1.outcomes, v2.outcomes) and int(v1.outcomes.sum()) > 0
