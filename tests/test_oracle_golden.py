"""CPU: the oracle (oracle/aem_oracle.c) is pinned against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  ORCA outputs inside the fixtures come from our own restatement behind the rvo2 stub
(parity unpinned at that boundary, see test_orca_properties.py)."""
import os

import numpy as np
import pytest

from aemcarl_b200 import weights as wts

DECISIONS = ["circle5", "circle5_mixed", "circle5_visible", "square10", "circle18"]


def test_action_table_bit_exact(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "action_rotate.npz"))
    assert np.array_equal(oracle.build_action_space(1.0), g["actions"])
    assert abs(float(g["gamma_bar"]) - 0.9791483623609768) < 1e-16
    assert np.allclose(g["speeds"], [0.12885124808584156, 0.2862305178902687, 0.47845399210662953,
                                     0.7132362736976232, 1.0], rtol=0, atol=1e-16)


def test_rotate_matches_reference(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "action_rotate.npz"))
    out = oracle.rotate(g["rotate_in"])
    assert np.abs(out - g["rotate_out"]).max() <= 4e-6          # FP32 trig of two libms


@pytest.mark.parametrize("tag", ["readme", "config"])
def test_occupancy_reward_matches_reference(oracle, golden_dir, actions, tag):
    g = np.load(os.path.join(golden_dir, "reward_map.npz"))
    ats, hts, pv, dv = g[tag + "_params"]
    worst, nonzero = 0.0, 0
    for t in range(len(g[tag + "_robot"])):
        n = int(g[tag + "_n"][t])
        hs = g[tag + "_humans"][t][:n]
        rp = g[tag + "_robot"][t]
        for ai in range(0, 81, 2):
            a = actions[ai]
            p = oracle.occupancy(hs, rp[0] + a[0] * ats, rp[1] + a[1] * ats, 0.3, hts, pv, dv)
            ref = g[tag + "_prob"][t][ai]
            worst = max(worst, abs(p - ref))
            nonzero += ref > 0
    assert worst <= 1e-13 and nonzero > 300


@pytest.mark.parametrize("tag", DECISIONS)
def test_full_decision_pipeline_matches_reference(oracle, golden_dir, actions, tag):
    """robot/human states recorded while the reference played episodes -> rewards, info, next states, network
    values, ACT step counts and argmax of the reference's own predict()"""
    g = np.load(os.path.join(golden_dir, "decisions_%s.npz" % tag))
    pb = float(g["ponder_bias"])
    W = wts.synthetic_weights(int(g["weight_seed"]), None if np.isnan(pb) else pb)
    params = oracle.default_params(robot_visible=int(g["robot_visible"]))
    nv = oracle.orca(g["humans"], g["robot"], params)
    rew, info, dmin, hnext = oracle.env_lookahead(g["robot"], g["humans"], nv, g["global_time"], actions, params)
    assert np.array_equal(hnext, g["humans_next"])
    assert np.array_equal(info, g["infos"])
    assert np.abs(rew - g["rewards"]).max() <= 1e-13
    vals, net, best, steps = oracle.lookahead(W, g["robot"], g["humans_next"], g["rewards"], actions, nthreads=8)
    rel = np.abs(net - g["net_values"]) / np.maximum(np.abs(g["net_values"]), 1e-2)
    assert rel.max() <= 2e-5, rel.max()
    assert np.array_equal(steps, g["act_steps"])
    assert np.array_equal(best, g["best"])
    assert np.abs(vals - g["values"]).max() <= 1e-6


@pytest.mark.parametrize("tag", ["circle5", "square10", "circle18"])
def test_act_halting_counts_match_reference(oracle, golden_dir, actions, tag):
    g = np.load(os.path.join(golden_dir, "values_%s.npz" % tag))
    seen = set()
    for wname in ("k1", "k2", "k3", "mixed"):
        pb = float(g["ponder_bias_" + wname])
        W = wts.synthetic_weights(int(g["seed_" + wname]), None if np.isnan(pb) else pb)
        R = g["robot"].shape[0]
        vals, net, best, steps = oracle.lookahead(W, g["robot"], g["humans_next"], np.zeros((R, 81)), actions,
                                                  gamma_bar=1.0, nthreads=8)
        ref = g["net_values_" + wname]
        assert (np.abs(net - ref) / np.maximum(np.abs(ref), 1e-2)).max() <= 2e-5
        assert np.array_equal(steps, g["act_steps_" + wname])
        seen |= set(np.unique(steps).tolist())
    assert seen == {1, 2, 3}


def test_episode_replay_matches_reference(oracle, golden_dir, actions):
    """Re-play the reference's recorded episodes with the oracle pipeline: same action at every step, same outcome."""
    g = np.load(os.path.join(golden_dir, "decisions_circle5.npz"))
    from aemcarl_b200 import scenarios
    W = wts.synthetic_weights(int(g["weight_seed"]), None)
    params = oracle.default_params()
    gbar = pow(0.9, 0.2)
    for case in g["cases"][:2]:
        ref_actions = g["actions_case%d" % case]
        outcome, nsteps = g["outcome_case%d" % case]
        rb, hs = scenarios.generate_case("test", int(case), 5, "circle_crossing")
        robot, humans, gt = rb[None].copy(), hs[None].copy(), np.zeros(1)
        mismatch, code = 0, -1
        for step in range(int(nsteps)):
            nv = oracle.orca(humans, robot, params)
            rew, info, dmin, hnext = oracle.env_lookahead(robot, humans, nv, gt, actions, params)
            vals, net, best, steps = oracle.lookahead(W, robot, hnext, rew, actions, gamma_bar=gbar, nthreads=8)
            if best[0] != ref_actions[step]:
                v = vals[0]
                assert abs(v[best[0]] - v[ref_actions[step]]) <= 1e-6 * max(abs(v[best[0]]), 1e-3)
                mismatch += 1
            chosen = np.array([ref_actions[step]], dtype=np.int32)
            code = info[0, chosen[0]]
            oracle.env_apply(robot, humans, nv, gt, actions, chosen, params)
            if code >= 2:
                break
        assert code == outcome and step + 1 == nsteps
        assert mismatch <= 1


@pytest.mark.parametrize("tag,rule,rand", [("mixed5", "mixed", False), ("mixed5_rand", "mixed", True),
                                           ("circle5_rand", "circle_crossing", True),
                                           ("square5_rand", "square_crossing", True)])
def test_mixed_rule_and_randomized_attributes_match_reference(golden_dir, tag, rule, rand):
    """crowd_sim.py:337-386 ('mixed': static / dynamic scenes with a drawn human count, dummy human for an empty static
    scene) and Agent.sample_random_attributes (agent.py:39-45): start, goal, radius and v_pref of every human of every
    recorded case are bit-identical to the unmodified reference."""
    from aemcarl_b200 import scenarios
    g = np.load(os.path.join(golden_dir, "scenarios.npz"))
    rows, count, hnum = g[tag + "_rows"], g[tag + "_count"], g[tag + "_human_num"]
    assert len(set(count.tolist())) > (3 if rule == "mixed" else 0)
    for case in range(len(count)):
        rb, hs = scenarios.generate_case("test", case, 5, rule, randomize_attributes=rand)
        assert hs.shape == (count[case], 8)
        got = hs[:, [0, 1, 5, 6, 4, 7]]
        assert np.array_equal(got, rows[case, :count[case]]), (tag, case)
        if rule == "mixed":        # an empty static scene keeps human_num = 0 but carries the parked dummy human
            assert hnum[case] == (0 if (count[case] == 1 and hs[0, 1] == -10 and hs[0, 0] == 0) else count[case])
    if rule == "mixed":
        with pytest.raises(ValueError):                     # batches of mixed scenes carry per-environment counts:
            scenarios.generate_pool("test", 0, 4, 5, "mixed")
        ncase = len(count)
        pr, ph, pc = scenarios.generate_mixed_pool("test", 0, ncase, randomize_attributes=rand)
        assert ph.shape == (ncase, 5, 8) and pc.dtype == np.int32 and np.array_equal(pc, count)
        for i in range(ncase):                              # padded pool rows = the reference's cases, bit for bit
            assert np.array_equal(ph[i, :pc[i]][:, [0, 1, 5, 6, 4, 7]], rows[i, :pc[i]])


def test_explorer_update_memory_targets_match_the_reference(golden_dir):
    """Explorer.update_memory of the drop-in mirror (crowd_nav/utils/explorer.py:114-151) against the pairs the UNMODIFIED
    reference pushed into its ReplayMemory for the same episode (oracle/make_golden.py gen_update_memory): imitation targets
    (discounted Monte-Carlo return) and RL targets (r + gamma_bar V_target(s')), rotated states included.  Runs on the CPU: the
    torch restatement of the value network carries the target forward there."""
    import torch
    from aemcarl_b200 import weights as wts
    from aemcarl_b200.crowd_nav.common.value_network import ValueNetworkTransformerAndGRU
    from aemcarl_b200.crowd_nav.policy.actenvcarl import rotate
    from aemcarl_b200.crowd_nav.utils.explorer import Explorer
    from aemcarl_b200.crowd_nav.utils.memory import ReplayMemory
    from aemcarl_b200.crowd_sim.envs.utils.state import FullState, JointState, ObservableState
    g = np.load(os.path.join(golden_dir, "update_memory.npz"))
    model = ValueNetworkTransformerAndGRU().eval()
    model.load_flat_weights(wts.synthetic_weights(int(g["weight_seed"])))

    class _Policy(object):                       # the two members update_memory uses of the target policy
        gamma = float(g["gamma"])

        def transform(self, state):
            rows = [list(state.self_state.as_tuple()) + list(h.as_tuple()) for h in state.human_states]
            return rotate(torch.tensor(rows, dtype=torch.float32), "holonomic")

    class _Robot(object):
        time_step, v_pref = float(g["time_step"]), float(g["v_pref"])

    joint = [JointState(FullState(*[float(v) for v in rb]), [ObservableState(*[float(v) for v in h]) for h in hs])
             for rb, hs in zip(g["robot"], g["humans_obs"])]
    rewards = [float(r) for r in g["rewards"]]
    for il, tag in ((True, "il"), (False, "rl")):
        memory = ReplayMemory(1000)
        ex = Explorer(None, _Robot(), torch.device("cpu"), memory, _Policy.gamma, target_policy=_Policy())
        ex.target_model = model
        states = joint if il else [_Policy().transform(s) for s in joint]
        ex.update_memory(states, [None] * len(states), rewards, imitation_learning=il)
        got_s = np.stack([m[0].numpy() for m in memory.memory])
        got_v = np.array([float(m[1]) for m in memory.memory])
        assert got_s.shape == g[tag + "_states"].shape
        assert np.abs(got_s - g[tag + "_states"]).max() < 5e-6
        assert np.abs(got_v - g[tag + "_values"]).max() < 5e-6, (tag, np.abs(got_v - g[tag + "_values"]).max())
    assert np.abs(g["il_values"]).max() > 0.5 and np.abs(g["rl_values"]).max() > 0.5


@pytest.mark.parametrize("optimizer", ["Adam", "SGD", "RMSprop"])
def test_trainer_step_matches_the_reference(golden_dir, optimizer):
    """One Trainer.optimize_batch step (crowd_nav/utils/trainer.py:22-86) of the drop-in mirror against the UNMODIFIED reference on the
    same memory (the batch is the whole memory, so the shuffle only permutes it): loss and the parameters after the update
    (every 8th parameter is recorded), for the three optimizers set_learning_rate offers.  CPU, the eager torch path -- the
    fused CUDA step is compared with this path in tests/test_gpu_train.py."""
    import torch
    from aemcarl_b200 import weights as wts
    from aemcarl_b200.crowd_nav.common.value_network import ValueNetworkTransformerAndGRU
    from aemcarl_b200.crowd_nav.utils.memory import ReplayMemory
    from aemcarl_b200.crowd_nav.utils.trainer import Trainer
    g = np.load(os.path.join(golden_dir, "update_memory.npz"))
    model = ValueNetworkTransformerAndGRU().eval()
    model.load_flat_weights(wts.synthetic_weights(int(g["weight_seed"])))
    memory = ReplayMemory(1000)
    for s, v in zip(g["rl_states"], g["rl_values"]):
        memory.push((torch.from_numpy(s.copy()), torch.tensor([float(v)])))
    tr = Trainer(model, memory, torch.device("cpu"), len(memory.memory))
    tr.set_learning_rate(0.001, optimizer)
    torch.manual_seed(0)
    loss = tr.optimize_batch(1)
    assert loss == pytest.approx(float(g["trainer_%s_loss" % optimizer]), rel=1e-5)
    after = model.flat_weights()[::8]
    ref = g["trainer_%s_params" % optimizer]
    before = wts.synthetic_weights(int(g["weight_seed"]))[::8]
    step = np.abs(ref - before).max()
    assert step > 1e-5                                                        # the update moved the parameters
    err = np.abs(after - ref)
    print("%s: largest update %.3e, largest deviation from the reference %.3e" % (optimizer, step, err.max()))
    if optimizer == "RMSprop":
        # the first RMSprop step is lr * g / (|g| sqrt(1 - alpha) + eps): where |g| is at rounding level its sign is noise
        assert (err > 1e-3 * step).mean() < 0.01 and err.max() <= 2.5 * step
    else:
        assert err.max() <= 1e-3 * step, (err.max(), step)
