"""GPU: the vectorised Explorer / on-device replay memory against the serial drop-in Explorer
(crowd_nav/utils/explorer.py:22-151 mirror, itself pinned to the reference's recorded episodes in test_gpu_api.py)
and against the oracle's rotate()."""
import numpy as np
import pytest
import torch

from aemcarl_b200.crowd_nav.utils.explorer import Explorer
from aemcarl_b200.crowd_nav.utils.memory import ReplayMemory
from aemcarl_b200.crowd_nav.utils.trainer import Trainer
from aemcarl_b200.crowd_nav.utils.vec_explorer import DeviceReplayMemory, VecExplorer
from test_gpu_api import setup

pytestmark = pytest.mark.gpu


def test_transform_matches_oracle_rotate(oracle):
    """aem_transform = MultiHumanRL.transform (multi_human_rl.py:423-438): bit-exact with the oracle's rotate()"""
    from aemcarl_b200 import scenarios
    from aemcarl_b200.engine import Engine
    eng = Engine(0)
    B, n = 64, 7
    robot, humans = scenarios.generate_pool("train", 11, B, n, "circle_crossing")
    rng = np.random.RandomState(3)
    robot[:, 2:4] = rng.uniform(-1, 1, (B, 2))            # moving robot, theta
    robot[:, 8] = rng.uniform(-3, 3, B)
    humans[:, :, 2:4] = rng.uniform(-1, 1, (B, n, 2))
    out = eng.transform(torch.from_numpy(robot).cuda(), torch.from_numpy(humans).cuda()).cpu().numpy()
    in14 = np.zeros((B, n, 14), dtype=np.float32)
    in14[:, :, :9] = robot[:, None, :].astype(np.float32)
    in14[:, :, 9:] = humans[:, :, :5].astype(np.float32)
    ref = oracle.rotate(in14.reshape(-1, 14)).reshape(B, n, 13)
    assert np.array_equal(out, ref)
    eng.close()


def _sorted_pairs(states, values):
    states = np.asarray(states, dtype=np.float64).reshape(len(values), -1)
    values = np.asarray(values, dtype=np.float64).reshape(-1)
    key = np.lexsort(np.concatenate([states.T[::-1], values[None]], axis=0))
    return states[key], values[key]


def test_vectorised_explorer_reproduces_the_serial_explorer():
    """Same policy, same target network, same 8 training cases, epsilon = 0: the (state, value) pairs pushed by the
    vectorised explorer are the pairs the serial one pushes (order aside), and the statistics agree."""
    env, robot, policy = setup()
    dev = torch.device("cuda:0")
    policy.set_phase("train")
    policy.set_epsilon(0.0)
    k = 8
    mem = ReplayMemory(100000)
    serial = Explorer(env, robot, dev, mem, policy.gamma, policy)
    serial.update_target_model(policy.get_model())
    env.case_counter["train"] = 0
    s_stats = serial.run_k_episodes(k, "train", update_memory=True)
    s_states = [m[0].cpu().numpy() for m in mem.memory]
    s_values = [m[1].item() for m in mem.memory]

    dmem = DeviceReplayMemory(100000, 5, dev)
    vec = VecExplorer(policy, 5, env.env_params(), dmem, policy.gamma, human_num=5)     # 5 envs < 8 cases: refills
    vec.update_target_model(policy.get_model())
    v_stats = vec.run_k_episodes(k, "train", update_memory=True, epsilon=0.0, first_case=0)
    for key in ("success", "collision", "timeout", "too_close"):
        assert v_stats[key] == s_stats[key], key
    assert v_stats["avg_nav_time"] == pytest.approx(s_stats["avg_nav_time"], abs=1e-9)
    assert np.allclose(sorted(v_stats["cumulative_rewards"]), sorted(s_stats["cumulative_rewards"]), atol=1e-9)
    assert len(dmem) == len(mem) == v_stats["pushed"] and len(mem) > 20
    a_s, a_v = _sorted_pairs(s_states, s_values)
    b_s, b_v = _sorted_pairs(dmem.states[:len(dmem)].cpu().numpy(), dmem.values[:len(dmem)].cpu().numpy())
    assert np.allclose(a_s, b_s, atol=2e-6), np.abs(a_s - b_s).max()
    assert np.allclose(a_v, b_v, atol=5e-6), np.abs(a_v - b_v).max()
    assert vec.case_counter["train"] == k


def test_epsilon_greedy_and_training_on_the_device_memory():
    env, robot, policy = setup()
    dev = torch.device("cuda:0")
    dmem = DeviceReplayMemory(5000, 5, dev)
    vec = VecExplorer(policy, 16, env.env_params(), dmem, policy.gamma, human_num=5, seed=1)
    vec.update_target_model(policy.get_model())
    greedy = vec.run_k_episodes(16, "train", update_memory=False, epsilon=0.0, first_case=100)
    noisy = vec.run_k_episodes(16, "train", update_memory=True, epsilon=0.5, first_case=100)
    assert noisy["env_steps"] != greedy["env_steps"] or noisy["cumulative_rewards"] != greedy["cumulative_rewards"]
    assert len(dmem) == noisy["pushed"] > 0
    trainer = Trainer(policy.get_model(), dmem, dev, batch_size=32)
    trainer.set_learning_rate(1e-3, "Adam")
    policy.get_model().train()
    l0 = trainer.optimize_batch(3)
    l1 = trainer.optimize_epoch(1)
    policy.get_model().eval()
    assert np.isfinite(l0) and np.isfinite(l1)


def test_train_entry_point_with_the_vectorised_rl_phase(tmp_path):
    """train.py --vectorised B: the imitation phase (ORCA teacher) and the RL waves both run B episodes in lock-step and
    fill the device memory, the optimizer runs on device-gathered mini-batches, the target network is refreshed across a
    wave boundary."""
    from aemcarl_b200.crowd_nav import train
    out = train.main(["--no_final_test", "--output_dir", str(tmp_path), "--il_episodes", "6", "--il_epochs", "2", "--train_episodes", "64",
                      "--vectorised", "32", "--updates_per_wave", "4", "--cuda_graph"])
    ex = out["explorer"]
    assert out["memory"] == len(ex.device_memory) > 50         # imitation + RL experience, all in the device memory
    assert ex.vec_stats["success"] + ex.vec_stats["collision"] + ex.vec_stats["timeout"] == 32
    assert ex.vec.case_counter["train"] > 0
    assert (tmp_path / "rl_model.pth").exists()
    assert all(torch.isfinite(p).all() for p in out["model"].parameters())


def test_train_entry_point_with_the_fused_optimisation_step(tmp_path):
    """train.py --fused_step: imitation epochs and RL updates run through aem_train_grad + aem_adam_step (csrc/train.cu);
    the policy's lookahead engine keeps seeing the updated weights, the checkpoint holds the trained parameters."""
    from aemcarl_b200.crowd_nav import train
    out = train.main(["--no_final_test", "--output_dir", str(tmp_path), "--il_episodes", "6", "--il_epochs", "2", "--train_episodes", "64",
                      "--vectorised", "32", "--updates_per_wave", "4", "--fused_step"])
    model = out["model"]
    assert all(torch.isfinite(p).all() for p in model.parameters())
    assert out["trainer"]._fused is not None and out["trainer"]._fused.step > 0
    sd = torch.load(tmp_path / "rl_model.pth", map_location="cpu")
    for k, v in model.state_dict().items():
        assert torch.equal(sd[k], v.cpu()), k
    ref = type(model)()
    ref.load_state_dict(sd)                                    # a reference-shaped module loads the checkpoint
    assert out["explorer"].vec_stats["success"] + out["explorer"].vec_stats["collision"] \
        + out["explorer"].vec_stats["timeout"] == 32


def test_vectorised_imitation_phase_reproduces_the_serial_teacher():
    """Imitation learning (train.py:171-204): the robot follows the ORCA teacher (safety_space 0.15), every step is
    stored with the discounted Monte-Carlo return.  Vectorised (robot appended to the humans as one more ORCA agent,
    aem_env_step_xy for the continuous action) against the serial drop-in Explorer on the same 6 cases."""
    from aemcarl_b200.crowd_nav.policy.policy_factory import policy_factory
    env, robot, policy = setup()
    dev = torch.device("cuda:0")
    teacher = policy_factory["orca"]()
    teacher.multiagent_training = policy.multiagent_training
    teacher.safety_space = 0.15
    robot.set_policy(teacher)
    k = 6
    mem = ReplayMemory(100000)
    serial = Explorer(env, robot, dev, mem, policy.gamma, target_policy=policy)
    env.case_counter["train"] = 0
    s_stats = serial.run_k_episodes(k, "train", update_memory=True, imitation_learning=True)
    s_states = [m[0].cpu().numpy() for m in mem.memory]
    s_values = [m[1].item() for m in mem.memory]
    params = env.env_params()
    params.orca_safety_space = 0.0          # the humans' own ORCA (env.env_params() reads it from the humans' policy)

    dmem = DeviceReplayMemory(100000, 5, dev)
    vec = VecExplorer(policy, 4, params, dmem, policy.gamma, human_num=5, il_safety_space=0.15)
    v_stats = vec.run_k_episodes(k, "train", update_memory=True, imitation_learning=True, first_case=0)
    for key in ("success", "collision", "timeout", "too_close"):
        assert v_stats[key] == s_stats[key], key
    assert v_stats["avg_nav_time"] == pytest.approx(s_stats["avg_nav_time"], abs=1e-9)
    assert len(dmem) == len(mem) > 100
    a_s, a_v = _sorted_pairs(s_states, s_values)
    b_s, b_v = _sorted_pairs(dmem.states[:len(dmem)].cpu().numpy(), dmem.values[:len(dmem)].cpu().numpy())
    assert np.allclose(a_s, b_s, atol=2e-5), np.abs(a_s - b_s).max()
    assert np.allclose(a_v, b_v, atol=2e-6), np.abs(a_v - b_v).max()


def test_cuda_graph_training_step_equals_the_eager_step():
    """the optimisation step replayed from CUDA graphs (control-flow-free ACT, Adam capturable) walks the
    same trajectory as the eager step: same losses and parameters after 6 steps on the same mini-batches, and the
    warm-up / capture leaves no trace."""
    import copy
    env, robot, policy = setup(seed=4, ponder_bias=-0.13)        # weights whose halting step varies with the batch
    dev = torch.device("cuda:0")
    model_a = policy.get_model()
    model_b = copy.deepcopy(model_a).to(dev)
    g = torch.Generator(device=dev).manual_seed(5)
    batches = [(torch.randn((100, 5, 13), device=dev, generator=g), torch.randn((100, 1), device=dev, generator=g))
               for _ in range(6)]
    losses = {}
    for name, model, graph in (("eager", model_a, False), ("graph", model_b, True)):
        tr = Trainer(model, None, dev, 100, cuda_graph=graph)
        tr.cuda_graph = True                      # both use Adam(capturable=True); only one replays graphs
        tr.set_learning_rate(1e-3, "Adam")
        tr.cuda_graph = graph
        losses[name] = [float(tr._step(x, v)) for x, v in batches]
    assert np.allclose(losses["eager"], losses["graph"], rtol=1e-5, atol=1e-7), (losses["eager"], losses["graph"])
    for pa, pb in zip(model_a.parameters(), model_b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-4, atol=1e-6)
    assert losses["graph"][0] != losses["graph"][-1]


def test_vectorised_explorer_phases_wave_sizes_and_rules():
    """VecExplorer on a policy whose action space has not been built yet: val / test waves (host-drawn, bit-exact scenarios),
    training waves smaller and larger than the batch, and the square rule with 10 humans (device-drawn scenarios); every
    case ends in exactly one outcome and only training waves feed the memory."""
    env, robot, policy = setup()
    dev = torch.device("cuda:0")
    policy.set_phase("train")
    policy.set_epsilon(0.1)
    mem = DeviceReplayMemory(100000, 5, dev)
    vec = VecExplorer(policy, 64, env.env_params(), mem, policy.gamma, human_num=5)
    vec.update_target_model(policy.get_model())
    before = 0
    for k, phase in ((100, "val"), (37, "test"), (5, "train"), (200, "train")):
        st = vec.run_k_episodes(k, phase, update_memory=(phase == "train"), epsilon=0.1)
        assert st["success"] + st["collision"] + st["timeout"] == k
        assert st["env_steps"] >= k
        assert (len(mem) > before) == (phase == "train")
        assert len(mem) - before == st["pushed"]
        before = len(mem)
    mem10 = DeviceReplayMemory(10000, 10, dev)
    vec10 = VecExplorer(policy, 32, env.env_params(), mem10, policy.gamma, human_num=10, rule="square_crossing")
    vec10.update_target_model(policy.get_model())
    st = vec10.run_k_episodes(48, "train", update_memory=True, epsilon=0.0)
    assert st["success"] + st["collision"] + st["timeout"] == 48 and len(mem10) == st["pushed"] > 0


def test_repacking_the_running_environments_changes_nothing():
    """When a wave's cases have run out, VecExplorer re-packs the environments that are still playing to the front of the
    batch and runs the kernels on them alone.  The wave must be the same with and without it -- also with epsilon-greedy
    exploration (the draws are made per original slot): outcomes, navigation times, returns and the pushed (state, value) pairs."""
    env, robot, policy = setup()
    dev = torch.device("cuda:0")
    policy.set_phase("train")
    policy.set_epsilon(0.2)
    out = {}
    for name, frac in (("packed", 0.75), ("plain", None)):
        mem = DeviceReplayMemory(100000, 5, dev)
        vec = VecExplorer(policy, 256, env.env_params(), mem, policy.gamma, human_num=5, seed=3)
        vec.compact_below = frac
        vec.update_target_model(policy.get_model())
        st = vec.run_k_episodes(256, "train", update_memory=True, epsilon=0.2, first_case=500)
        out[name] = (st, mem.states[:len(mem)].cpu().numpy(), mem.values[:len(mem)].cpu().numpy())
    a, b = out["packed"], out["plain"]
    assert np.array_equal(a[0]["codes"], b[0]["codes"])
    for key in ("success", "collision", "timeout", "too_close", "env_steps", "pushed", "avg_nav_time"):
        assert a[0][key] == b[0][key], key
    assert np.array_equal(np.asarray(a[0]["cumulative_rewards"]), np.asarray(b[0]["cumulative_rewards"]))
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[0]["timeout"] > 0 and a[0]["success"] + a[0]["collision"] > 0      # the wave really had a long tail


def test_target_network_in_train_mode_is_the_references_stochastic_target():
    """target_train_mode: V_target of explorer.py:132-139 from the torch module with dropout live (the reference never calls
    eval()).  With p = 0 the torch path (per-sample halting) must give the targets of the eval()-mode kernel path; with the
    reference's p = 0.1 the targets scatter around them and differ from call to call."""
    env, robot, policy = setup()
    dev = torch.device("cuda:0")
    policy.set_phase("train")

    def targets(train_mode, p=None, seed=0):
        dmem = DeviceReplayMemory(100000, 5, dev)
        vec = VecExplorer(policy, 8, env.env_params(), dmem, policy.gamma, human_num=5, target_train_mode=train_mode)
        vec.update_target_model(policy.get_model())
        if p is not None:
            for m in vec.target_model.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = p
                if hasattr(m, "dropout") and isinstance(getattr(m, "dropout"), float):
                    m.dropout = p                      # nn.MultiheadAttention keeps its own probability
        torch.manual_seed(seed)
        vec.run_k_episodes(8, "train", update_memory=True, epsilon=0.0, first_case=0)
        st = dmem.states[:len(dmem)].cpu().numpy().astype(np.float64).reshape(len(dmem), -1)
        va = dmem.values[:len(dmem)].cpu().numpy().astype(np.float64).reshape(-1)
        key = np.lexsort(st.T[::-1])                     # by the state alone: the same order whatever the targets are
        return st[key], va[key]

    s0, v0 = targets(False)
    s1, v1 = targets(True, p=0.0)
    assert np.array_equal(s0, s1) and np.abs(v0 - v1).max() <= 5e-5
    s2, v2 = targets(True, seed=1)
    s3, v3 = targets(True, seed=2)
    assert np.array_equal(s0, s2) and np.isfinite(v2).all()
    d = np.abs(v2 - v0)
    assert 1e-4 < d.max() < 0.5 and not np.array_equal(v2, v3)
