"""GPU: per-environment human counts -- the 'mixed' scenario rule of CrowdSim.reset (crowd_sim/envs/crowd_sim.py:337-386:
static / dynamic scenes with 1..5 humans) on the batched path.  One batch holds environments with different human counts
(padded rows + aem_set_active_counts); every kernel must give, for each environment, exactly what the oracle gives for that
environment alone with its own number of humans."""
import numpy as np
import pytest
import torch

from aemcarl_b200 import scenarios, weights as wts
from aemcarl_b200.crowd_nav.utils.explorer import run_vectorised_test
from aemcarl_b200.engine import Engine, make_env_params
from aemcarl_b200.vec_env import VecCrowdSim
from test_gpu_episodes import LockstepProbe, oracle_episodes

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_mixed_pool_matches_the_reference_scenarios(golden_dir):
    """the padded pool rows are the bit-exact host scenarios (tests/golden/scenarios.npz holds the reference's mixed cases)"""
    robot, humans, counts = scenarios.generate_mixed_pool("test", 0, 40)
    assert set(counts.tolist()) <= {1, 2, 3, 4, 5} and len(set(counts.tolist())) >= 3
    for i in range(40):
        rb, hs = scenarios.generate_case("test", i, 5, "mixed")
        assert hs.shape[0] == counts[i] and np.array_equal(humans[i, :counts[i]], hs) and np.array_equal(robot[i], rb)


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
@pytest.mark.parametrize("visible", [False, True])
def test_kernels_with_active_counts_equal_the_oracle_per_environment(oracle, actions, precision, visible):
    B = 96
    rng = np.random.RandomState(17)
    robot, humans, counts = scenarios.generate_mixed_pool("train", 50, B)
    humans[:, :, 2:4] = rng.uniform(-0.7, 0.7, (B, 5, 2))                     # moving humans, robot somewhere on its way
    robot[:, 0:2] += rng.uniform(-1, 3, (B, 2))
    for b in range(B):
        humans[b, counts[b]:, 2:4] = 0.0
    gt = rng.uniform(0, 15, B)
    W = wts.synthetic_weights(2, -0.05)
    eng = Engine(0)
    eng.set_action_space(actions)
    eng.set_env_params(make_env_params(robot_visible=visible))
    eng.load_weights(W)
    eng.set_precision(precision)
    gbar = pow(0.9, 0.2)
    nact = dev(counts)
    eng.set_active_counts(nact, None)
    r_d, h_d, t_d = dev(robot), dev(humans), dev(gt)
    nv = eng.orca(h_d, r_d)
    rew, info, dmin, hnext = eng.env_lookahead(r_d, h_d, nv, t_d)
    values, net, best, steps = eng.lookahead(r_d, hnext, rew, gbar)
    chosen = best.clone()
    r2, h2, t2 = r_d.clone(), h_d.clone(), t_d.clone()
    eng.env_apply(r2, h2, nv, t2, chosen, rew, info, dmin)
    nv, rew, info, hnext, values, net, best, steps, r2, h2 = [x.cpu().numpy() for x in
                                                            (nv, rew, info, hnext, values, net, best, steps, r2, h2)]
    p = oracle.default_params(robot_visible=int(visible))
    for k in sorted(set(counts.tolist())):
        idx = np.nonzero(counts == k)[0]
        hk = np.ascontiguousarray(humans[idx, :k])
        rk, tk = robot[idx].copy(), gt[idx].copy()
        onv = oracle.orca(hk, rk, p)
        assert np.array_equal(nv[idx, :k], onv), k
        assert (nv[idx, k:] == 0).all()
        orew, oinfo, odmin, ohnext = oracle.env_lookahead(rk, hk, onv, tk, actions, p)
        assert np.array_equal(hnext[idx, :k], ohnext) and np.array_equal(info[idx], oinfo)
        assert np.abs(rew[idx] - orew).max() <= 1e-12
        ov, onet, obest, osteps = oracle.lookahead(W, rk, ohnext, orew, actions, gamma_bar=gbar, nthreads=8)
        same = steps[idx] == osteps
        assert same.mean() > 0.999
        rel = np.abs(net[idx] - onet) / np.maximum(np.abs(onet), 1e-2)
        assert rel[same].max() <= 1e-4, (k, rel[same].max())
        for j in np.nonzero(best[idx] != obest)[0]:
            v = ov[j]
            assert abs(v[best[idx][j]] - v[obest[j]]) <= 1e-4 * max(abs(v[obest[j]]), 1e-3)
        hk2 = hk.copy()
        oracle.env_apply(rk, hk2, onv, tk, actions, best[idx].astype(np.int32), p)
        assert np.array_equal(r2[idx], rk) and np.array_equal(h2[idx, :k], hk2)
        assert np.array_equal(h2[idx, k:], humans[idx, k:])                   # padding rows untouched
    eng.set_active_counts(None, None)
    eng.close()


def test_mixed_episodes_outcomes_match_the_oracle(oracle, actions):
    """120 'mixed' test cases played in ONE batch (VecCrowdSim rule 'mixed') end like the oracle's episodes of the same cases,
    played per human count; decisions are probed in lock-step like tests/test_gpu_episodes.py"""
    B = 120
    W = wts.synthetic_weights(1)
    eng = Engine(0)
    eng.set_action_space(actions)
    eng.set_env_params(make_env_params())
    eng.load_weights(W)
    eng.set_precision("fp16x3")
    venv = VecCrowdSim(eng, B, 5, "mixed", "test", first_case=0, pool_size=B)
    out = run_vectorised_test(venv, B)
    eng.set_active_counts(None, None)
    robot, humans, counts = scenarios.generate_mixed_pool("test", 0, B)
    codes = np.full(B, -1)
    excused = {}
    for k in sorted(set(counts.tolist())):
        idx = np.nonzero(counts == k)[0]
        probe = LockstepProbe(eng, oracle, W, actions, pow(0.9, 0.2))
        ck, _ = oracle_episodes(oracle, W, actions, robot[idx].copy(), np.ascontiguousarray(humans[idx, :k]), probe=probe)
        assert not probe.violations, (k, probe.violations[:5])
        codes[idx] = ck
        excused.update({int(idx[c]): v for c, v in probe.excused.items()})
    differ = [int(c) for c in np.nonzero(out["codes"] != codes)[0] if int(c) not in excused]
    assert not differ, differ[:10]
    assert out["success"] + out["collision"] + out["timeout"] == B
    print("mixed scenes: counts %s, outcomes (success, collision, timeout) %d %d %d, excused %d"
          % (np.bincount(counts, minlength=6).tolist(), out["success"], out["collision"], out["timeout"], len(excused)))
    eng.close()


@pytest.mark.parametrize("randomize", [False, True])
def test_device_mixed_scenarios_replay_the_host_draw(randomize):
    """aem_generate_mixed_scenarios against scenarios.generate_mixed_pool (which reproduces the reference's recorded mixed
    cases bit for bit): the same static / dynamic decision and human count for every case, static and square-crossing
    humans bit-exact, circle-crossing ones (the first two of a dynamic scene) within 2 ulp of the 4 m radius"""
    eng = Engine(0)
    for phase, first in (("train", 77), ("val", 0), ("test", 400)):
        r_h, h_h, c_h = scenarios.generate_mixed_pool(phase, first, 512, randomize_attributes=randomize)
        r_d, h_d, c_d = eng.generate_mixed_scenarios(phase, first, 512, randomize_attributes=randomize)
        r_d, h_d, c_d = r_d.cpu().numpy(), h_d.cpu().numpy(), c_d.cpu().numpy()
        assert np.array_equal(c_d, c_h) and np.array_equal(r_d, r_h)
        assert len(set(c_h.tolist())) >= 4
        static = np.array([(h_h[i, :c_h[i], 0:2] == h_h[i, :c_h[i], 5:7]).all() for i in range(512)])
        assert 0.1 < static.mean() < 0.3                                    # goal = start: the static scenes
        assert np.array_equal(h_d[static], h_h[static])
        assert np.array_equal(h_d[:, 2:], h_h[:, 2:])                        # square-crossing humans and padding rows
        assert np.abs(h_d - h_h).max() <= 2 * np.spacing(4.0)
        assert np.array_equal(h_d[:, :, [2, 3, 4, 7]], h_h[:, :, [2, 3, 4, 7]])


def test_vec_env_runs_device_generated_mixed_scenes():
    """VecCrowdSim(rule='mixed', device_scenarios=True): the pool and its counts never touch the host; the outcomes equal the
    host-generated pool's wherever no circle-crossing human's last-ulp difference matters (all of them here)"""
    W = wts.synthetic_weights(2, -0.05)
    res = []
    for dev_scn in (False, True):
        eng = Engine(0)
        eng.set_action_space(np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden",
                                                               "action_rotate.npz"))["actions"])
        eng.set_env_params(make_env_params())
        eng.load_weights(W)
        eng.set_precision("fp16x3")
        env = VecCrowdSim(eng, 64, 5, rule="mixed", phase="test", pool_size=128, device_scenarios=dev_scn)
        for _ in range(120):
            env.decide_and_step()
        torch.cuda.synchronize()
        res.append(env.outcomes.cpu().numpy().copy())
    assert res[0].sum() >= 64 and np.array_equal(res[0], res[1])
