"""GPU: episode outcomes on seeded test cases (BASELINE.json configs[0]: circle_crossing, 5 humans, test seeds
1000..1499, plus square_crossing with 10 humans, circle_crossing with 18 humans and the visible-robot mode) -- the
batched GPU pipeline against the CPU oracle pipeline on identical scenarios.

north_star: "episode outcomes (success / collision / timeout) identical on the 500 test cases".  The test asserts
IDENTITY, case by case.  The only admitted exception is the documented one: while the oracle plays its episodes the
GPU pipeline evaluates every one of the oracle's states as well (lock-step probe), and a case may end differently only
if one of its decisions was a near-tie (top-2 gap <= 1e-4 |value|, reference explorer.py:44-98 picks by strict `>`)
or an ACT halting flip within the documented margin; every such decision is listed and checked.
"""
import numpy as np
import pytest
import torch

from aemcarl_b200 import scenarios, weights as wts
from aemcarl_b200.crowd_nav.utils.explorer import run_vectorised_test
from aemcarl_b200.engine import Engine, make_env_params
from aemcarl_b200.vec_env import VecCrowdSim

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def oracle_episodes(po, W, actions, robot, humans, max_steps=110, params=None, probe=None):
    """Explorer.run_k_episodes('test') for all cases at once with the oracle: returns outcome code and length per case.
    `probe(idx, robot, humans, time, values, best, steps)` sees every decision of the oracle before it is applied."""
    B = robot.shape[0]
    p = params if params is not None else po.default_params()
    gbar = pow(0.9, 0.2)
    gt = np.zeros(B)
    codes = np.full(B, -1)
    length = np.zeros(B, dtype=np.int64)
    alive = np.ones(B, dtype=bool)
    for step in range(max_steps):
        idx = np.nonzero(alive)[0]
        if idx.size == 0:
            break
        r, h, t = robot[idx].copy(), humans[idx].copy(), gt[idx].copy()
        nv = po.orca(h, r, p, 16)
        rew, info, dmin, hnext = po.env_lookahead(r, h, nv, t, actions, p, 16)
        vals, net, best, steps = po.lookahead(W, r, hnext, rew, actions, gamma_bar=gbar, nthreads=16)
        if probe is not None:
            probe(idx, r, h, t, hnext, vals, best, steps)
        po.env_apply(r, h, nv, t, actions, best, p)
        robot[idx], humans[idx], gt[idx] = r, h, t
        c = info[np.arange(idx.size), best]
        length[idx] += 1
        done = c >= 2
        codes[idx[done]] = c[done]
        alive[idx[done]] = False
    return codes, length


class LockstepProbe(object):
    """The GPU pipeline on the oracle's own states: every decision must agree, or be a documented near-tie / ACT flip."""

    def __init__(self, eng, po, W, actions, gbar):
        self.eng, self.po, self.W, self.actions, self.gbar = eng, po, W, actions, gbar
        self.decisions = 0
        self.excused = {}          # case -> list of (kind, detail) of its non-identical decisions
        self.violations = []

    def __call__(self, idx, r, h, t, hnext, vals, best, steps):
        e = self.eng
        rd, hd, td = _dev(r), _dev(h), _dev(t)
        nv = e.orca(hd, rd)
        rew_g, info_g, dmin_g, hnext_g = e.env_lookahead(rd, hd, nv, td)
        v_g, net_g, best_g, steps_g = e.lookahead(rd, hnext_g, rew_g, self.gbar)
        bg, sg = best_g.cpu().numpy(), steps_g.cpu().numpy()
        self.decisions += idx.size
        assert np.array_equal(hnext_g.cpu().numpy(), hnext)
        for j in np.nonzero(bg != best)[0]:
            v = vals[j]
            gap = abs(v[bg[j]] - v[best[j]])
            flips = np.nonzero(sg[j] != steps[j])[0]
            if gap <= 1e-4 * max(abs(v[best[j]]), 1e-3):
                self.excused.setdefault(int(idx[j]), []).append(("near-tie", float(gap)))
            elif flips.size and all(self.po.act_margin(self.W, self.po.rotated_state(r[j], hnext[j], self.actions[a])) <= 1e-4
                                    for a in flips):
                self.excused.setdefault(int(idx[j]), []).append(("act-margin", flips.tolist()))
            else:
                self.violations.append((int(idx[j]), int(best[j]), int(bg[j]), float(gap)))


def run_case_set(oracle, actions, W, precision, B, n, rule, visible=False):
    eng = Engine(0)
    eng.set_action_space(actions)
    eng.set_env_params(make_env_params(robot_visible=visible))
    eng.load_weights(W)
    eng.set_precision(precision)
    venv = VecCrowdSim(eng, B, n, rule, "test", first_case=0, pool_size=B)
    out = run_vectorised_test(venv, B)
    robot, humans = scenarios.generate_pool("test", 0, B, n, rule)
    probe = LockstepProbe(eng, oracle, W, actions, pow(0.9, 0.2))
    codes, length = oracle_episodes(oracle, W, actions, robot, humans, params=oracle.default_params(robot_visible=int(visible)),
                                    probe=probe)
    eng.close()
    assert (codes >= 2).all()
    assert not probe.violations, probe.violations[:10]
    differ = np.nonzero(out["codes"] != codes)[0]
    unexplained = [int(c) for c in differ if int(c) not in probe.excused]
    assert not unexplained, ("cases with a different outcome and no near-tie decision", unexplained[:10])
    assert out["success"] + out["collision"] + out["timeout"] == B
    if not probe.excused:
        assert out["env_steps"] == int(length.sum())
    print("[%s %s n=%d visible=%d] outcomes (success, collision, timeout) %d %d %d; %d of %d cases identical; "
          "%d decisions probed in lock-step, %d excused as near-tie / ACT margin: %s"
          % (precision, rule, n, int(visible), out["success"], out["collision"], out["timeout"], B - differ.size, B,
             probe.decisions, sum(len(v) for v in probe.excused.values()), dict(list(probe.excused.items())[:5])))
    return out, codes, probe


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_500_test_cases_outcomes_match_oracle(oracle, actions, precision):
    run_case_set(oracle, actions, wts.synthetic_weights(1), precision, 500, 5, "circle_crossing")


@pytest.mark.parametrize("n,rule,visible", [(10, "square_crossing", False), (18, "circle_crossing", False),
                                            (5, "circle_crossing", True)])
def test_100_cases_other_scenes_match_oracle(oracle, actions, n, rule, visible):
    """BASELINE configs[2] / configs[3] shapes (square crossing with 10 humans, 18-human circle: more than 10 ORCA
    neighbours) and the visible-robot mode (test.py --visible): 100 seeded test cases each, FP16x3 kernel."""
    run_case_set(oracle, actions, wts.synthetic_weights(1), "fp16x3", 100, n, rule, visible)
