"""GPU: the drop-in API (crowd_sim CrowdSim.reset/step/onestep_lookahead, crowd_nav ACTENVCARL.predict, Explorer)
against episodes recorded from the UNMODIFIED reference and against the oracle."""
import os

import numpy as np
import pytest
import torch

from aemcarl_b200 import weights as wts
from aemcarl_b200.crowd_nav.default_configs import env_config, policy_config
from aemcarl_b200.crowd_nav.policy.policy_factory import policy_factory
from aemcarl_b200.crowd_nav.utils.explorer import Explorer, run_vectorised_test
from aemcarl_b200.crowd_nav.utils.memory import ReplayMemory
from aemcarl_b200.crowd_sim import make
from aemcarl_b200.crowd_sim.envs.utils.action import ActionXY
from aemcarl_b200.crowd_sim.envs.utils.info import Collision, Danger, Nothing, ReachGoal, Timeout
from aemcarl_b200.crowd_sim.envs.utils.robot import Robot
from aemcarl_b200.crowd_sim.envs.utils.state import FullState, JointState, ObservableState

pytestmark = pytest.mark.gpu

INFO_CLS = {0: Nothing, 1: Danger, 2: ReachGoal, 3: Collision, 4: Timeout}


def setup(n=5, sim="circle_crossing", seed=1, ponder_bias=None, visible=False):
    cfg = env_config(sim__human_num=n)
    env = make("CrowdSim-v0")
    env.configure(cfg)
    env.test_sim = sim
    robot = Robot(cfg, "robot")
    robot.visible = visible
    policy = policy_factory["actenvcarl"]()
    policy.configure(policy_config())
    policy.get_model().load_flat_weights(wts.synthetic_weights(seed, ponder_bias))
    policy.set_phase("test")
    policy.set_device(torch.device("cuda:0"))
    robot.set_policy(policy)
    env.set_robot(robot)
    policy.set_env(env)
    return env, robot, policy


@pytest.mark.parametrize("tag,n,sim,visible,max_steps", [("circle5", 5, "circle_crossing", False, 97),
                                                         ("circle5_visible", 5, "circle_crossing", True, 40),
                                                         ("square10", 10, "square_crossing", False, 36),
                                                         ("circle18", 18, "circle_crossing", False, 22)])
def test_episode_matches_reference(golden_dir, tag, n, sim, visible, max_steps):
    """Explorer's inner loop (explorer.py:46-56) with our env + policy reproduces the reference's recorded episode:
    same action index at every step (near-ties excepted), same rewards, same final outcome."""
    g = np.load(os.path.join(golden_dir, "decisions_%s.npz" % tag))
    env, robot, policy = setup(n, sim, int(g["weight_seed"]), None, visible)
    case = int(g["cases"][0])
    ref_actions, ref_rewards = g["actions_case%d" % case], g["rewards_case%d" % case]
    outcome, nsteps = g["outcome_case%d" % case]
    ob = env.reset("test", case)
    done, step, near_ties = False, 0, 0
    while not done and step < max_steps:
        action = robot.act(ob)
        idx = policy.action_space.index(action)
        if idx != ref_actions[step]:
            v = np.array(policy.action_values)
            assert abs(v[idx] - v[ref_actions[step]]) <= 1e-4 * max(abs(v[idx]), 1e-3)
            near_ties += 1
            action = policy.action_space[ref_actions[step]]       # stay on the reference trajectory
        assert len(policy.action_values) == 81
        ob, reward, done, info = env.step(action)
        assert abs(reward - ref_rewards[step]) <= 1e-12
        step += 1
    assert near_ties <= 2
    assert step == nsteps
    if outcome >= 0:
        assert done and isinstance(info, INFO_CLS[int(outcome)])
        assert abs(env.global_time - float(g["time_case%d" % case])) < 1e-9
    assert sum(policy.get_act_statistic_info()) == step


def test_step_with_arbitrary_action_and_lookahead_vs_oracle(oracle):
    env, robot, policy = setup()
    env.reset("test", 3)
    rng = np.random.RandomState(0)
    p = oracle.default_params()
    for it in range(12):
        a = ActionXY(*rng.uniform(-0.7, 0.7, 2))
        r, h, t = env.pack_state()
        nv = oracle.orca(h, r, p)
        rew, info, dmin, hnext = oracle.env_lookahead(r, h, nv, t, np.array([[a.vx, a.vy]]), p)
        ob_l, reward_l, done_l, info_l = env.onestep_lookahead(a)
        assert np.array_equal(np.array([o.as_tuple() for o in ob_l]), hnext[0])
        assert abs(reward_l - rew[0, 0]) <= 1e-12 and isinstance(info_l, INFO_CLS[int(info[0, 0])])
        r0 = env.pack_state()
        assert np.array_equal(r0[0], r) and np.array_equal(r0[1], h)          # lookahead does not mutate
        ob, reward, done, inf = env.step(a)
        oracle.env_apply(r, h, nv, t, np.array([[a.vx, a.vy]]), np.zeros(1, dtype=np.int32), p)
        r1 = env.pack_state()
        assert np.array_equal(r1[0], r) and np.array_equal(r1[1], h) and abs(r1[2][0] - t[0]) < 1e-12
        assert reward == reward_l or abs(reward - reward_l) < 1e-15
        if done:
            break


def test_orca_policy_predict_vs_oracle(oracle):
    from aemcarl_b200.crowd_sim.envs.policy.orca import ORCA
    pol = ORCA()
    pol.time_step = 0.2
    pol.safety_space = 0.15                                     # IL teacher setting (train.py:186-195)
    me = FullState(0.0, 0.0, 0.5, 0.1, 0.3, 3.0, 2.0, 1.0, 0.0)
    others = [ObservableState(1.0, 0.3, -0.5, 0.0, 0.3), ObservableState(-0.5, 1.0, 0.2, -0.4, 0.35)]
    act = pol.predict(JointState(me, others))
    ref = oracle.orca_agent([0, 0], [0.5, 0.1], np.float32(0.3 + 0.01 + 0.15), 1.0,
                            np.array([3.0, 2.0]) / np.hypot(3.0, 2.0),
                            [[o.px, o.py, o.vx, o.vy, np.float32(o.radius + 0.01 + 0.15)] for o in others])
    assert (np.float32(act.vx), np.float32(act.vy)) == (ref[0], ref[1])


def test_vec_env_tracks_oracle(oracle, actions):
    from aemcarl_b200.engine import Engine, make_env_params
    from aemcarl_b200.vec_env import VecCrowdSim
    eng = Engine(0)
    eng.set_action_space(actions)
    eng.set_env_params(make_env_params())
    W = wts.synthetic_weights(1)
    eng.load_weights(W)
    B, n = 24, 5
    env = VecCrowdSim(eng, B, n, "circle_crossing", "test", first_case=0, pool_size=B)
    robot, humans, gt = env.state_numpy()
    p = oracle.default_params()
    gbar = env.gamma_bar()
    assert abs(gbar - 0.9791483623609768) < 1e-15
    for it in range(12):
        nv = oracle.orca(humans, robot, p)
        rew, info, dmin, hnext = oracle.env_lookahead(robot, humans, nv, gt, actions, p)
        vals, net, best, steps = oracle.lookahead(W, robot, hnext, rew, actions, gamma_bar=gbar, nthreads=8)
        chosen = env.lookahead().clone()
        c = chosen.cpu().numpy()
        for b in np.nonzero(c != best)[0]:
            assert abs(vals[b, c[b]] - vals[b, best[b]]) <= 1e-4 * max(abs(vals[b, best[b]]), 1e-3)
        reward, done, inf, dm = env.apply(chosen, auto_reset=False)
        oracle.env_apply(robot, humans, nv, gt, actions, c, p)
        r2, h2, t2 = env.state_numpy()
        assert np.array_equal(r2, robot) and np.array_equal(h2, humans) and np.allclose(t2, gt, atol=1e-12)
        assert np.array_equal(inf.cpu().numpy(), info[np.arange(B), c])
    eng.close()


def test_explorer_and_vectorised_evaluation_agree():
    env, robot, policy = setup()
    explorer = Explorer(env, robot, torch.device("cuda:0"), gamma=0.9)
    env.case_counter["test"] = 0
    stats = explorer.run_k_episodes(4, "test")
    assert stats["success"] + stats["collision"] + stats["timeout"] == 4
    from aemcarl_b200.vec_env import VecCrowdSim
    eng = policy.engine()
    eng.set_env_params(env.env_params())
    venv = VecCrowdSim(eng, 4, 5, "circle_crossing", "test", first_case=0, pool_size=4)
    out = run_vectorised_test(venv, 4)
    assert (out["success"], out["collision"], out["timeout"]) == (stats["success"], stats["collision"],
                                                                  stats["timeout"])


def test_update_memory_target_through_kernel_matches_torch():
    env, robot, policy = setup()
    policy.set_phase("train")
    policy.set_epsilon(0.0)
    mem = ReplayMemory(1000)
    explorer = Explorer(env, robot, torch.device("cuda:0"), mem, 0.9, policy)
    explorer.update_target_model(policy.get_model())
    ob = env.reset("train", 5)
    states, rewards = [], []
    for _ in range(4):
        a = robot.act(ob)
        ob, r, done, info = env.step(a)
        states.append(policy.last_state)
        rewards.append(r)
    assert states[0].shape == (5, 13) and states[0].is_cuda
    explorer.update_memory(states, [None] * 4, rewards, imitation_learning=False)
    assert len(mem) == 4
    s1 = states[1].unsqueeze(0)
    with torch.no_grad():
        v_kernel, _ = explorer.target_model(s1)
        v_torch, _ = explorer.target_model.forward_torch(s1)
    assert abs(v_kernel.item() - v_torch.item()) <= 2e-6 * max(abs(v_torch.item()), 1e-2) + 2e-7
    gbar = pow(0.9, 0.2)
    assert abs(mem[0][1].item() - (rewards[0] + gbar * v_kernel.item())) < 1e-6
    assert mem[3][1].item() == pytest.approx(rewards[3])


def test_train_entry_point_il_and_rl(tmp_path):
    """train.py flow end to end at a tiny size: IL episodes with the ORCA teacher fill the memory, the IL epochs reduce
    the loss, RL episodes run with epsilon-greedy predict, checkpoints carry the reference's state_dict keys."""
    from aemcarl_b200.crowd_nav import train
    out = train.main(["--no_final_test", "--output_dir", str(tmp_path), "--il_episodes", "12", "--il_epochs", "8", "--train_episodes", "2"])
    assert out["memory"] > 100
    sd = torch.load(str(tmp_path / "il_model.pth"), map_location="cpu")
    assert "in_mlp.rnn_cell.convz.0.weight" in sd and "encoder_layer.linear1.weight" in sd and len(sd) == 66
    assert (tmp_path / "rl_model.pth").exists()
    # a freshly configured policy accepts the checkpoint (test.py:105) and the trained value head is finite
    policy = policy_factory["actenvcarl"]()
    policy.configure(policy_config())
    policy.get_model().load_state_dict(sd)
    assert np.isfinite(out["loss"]) and out["loss"] < 1.0


def test_mixed_rule_episodes_run_with_a_varying_number_of_humans(golden_dir):
    """'mixed' scenes (crowd_sim.py:337-386) through the drop-in CrowdSim: the human count of every case equals the
    reference's and the episodes run to an outcome on the CUDA path (n changes from episode to episode)."""
    g = np.load(os.path.join(golden_dir, "scenarios.npz"))
    env, robot, policy = setup(5, "mixed")
    seen = set()
    for case in (2, 3, 4, 12):
        ob = env.reset("test", case)
        assert len(env.humans) == len(ob) == g["mixed5_count"][case] and env.human_num == g["mixed5_human_num"][case]
        seen.add(len(ob))
        done, steps = False, 0
        while not done and steps < 130:
            ob, reward, done, info = env.step(robot.act(ob))
            steps += 1
        assert done and isinstance(info, (ReachGoal, Collision, Timeout))
    assert len(seen) >= 3


@pytest.mark.parametrize("rule,n,randomize", [("square_crossing", 10, False), ("square_crossing", 5, True),
                                               ("circle_crossing", 5, False), ("circle_crossing", 5, True),
                                               ("circle_crossing", 10, False)])
def test_device_scenario_generation_replays_the_seeded_stream(rule, n, randomize):
    """aem_generate_scenarios against the host generator, which reproduces 208 cases recorded from the
    reference bit for bit (tests/test_oracle_golden.py): MT19937 stream, draw order and rejection sampling identical --
    square rule bit-exact; circle rule within 2 ulp of the 4 m radius (libdevice cos / sin against numpy's)."""
    from aemcarl_b200 import scenarios
    from aemcarl_b200.engine import get_engine
    eng = get_engine(0)
    for phase, first in (("train", 123), ("val", 0), ("test", 37)):
        r_h, h_h = scenarios.generate_pool(phase, first, 256, n, rule, randomize_attributes=randomize)
        r_d, h_d = eng.generate_scenarios(phase, first, 256, n, rule, randomize_attributes=randomize)
        r_d, h_d = r_d.cpu().numpy(), h_d.cpu().numpy()
        assert np.array_equal(r_d, r_h)
        if rule == "square_crossing":
            assert np.array_equal(h_d, h_h)
        else:
            assert np.abs(h_d - h_h).max() <= 2 * np.spacing(4.0)
            assert np.array_equal(h_d[:, :, [2, 3, 4, 7]], h_h[:, :, [2, 3, 4, 7]])


def test_device_scenario_generation_gives_up_on_an_infeasible_draw():
    """18 humans with randomised radii do not fit on the 4 m circle: the reference's rejection loop never returns; the kernel
    gives the case up (NaN) instead of hanging the GPU"""
    from aemcarl_b200.engine import get_engine
    r, h = get_engine(0).generate_scenarios("train", 0, 8, 30, "circle_crossing", randomize_attributes=True)
    torch.cuda.synchronize()
    assert torch.isnan(h).any()
