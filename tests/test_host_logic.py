"""CPU: host-side mirror of the reference interface -- scenario generation, configs, value types, the nn.Module
state_dict surface, replay memory, and the no-fallback rule."""
import os

import numpy as np
import pytest
import torch

from aemcarl_b200 import scenarios, weights as wts
from aemcarl_b200.crowd_nav.default_configs import env_config, policy_config, train_config
from aemcarl_b200.crowd_nav.policy.policy_factory import policy_factory
from aemcarl_b200.crowd_nav.utils.memory import ReplayMemory
from aemcarl_b200.crowd_sim import make
from aemcarl_b200.crowd_sim.envs.utils.action import ActionXY
from aemcarl_b200.crowd_sim.envs.utils.robot import Robot
from aemcarl_b200.crowd_sim.envs.utils.state import FullState, JointState, ObservableState


@pytest.mark.parametrize("tag,n,rule", [("circle5", 5, "circle_crossing"), ("square10", 10, "square_crossing"),
                                        ("circle18", 18, "circle_crossing")])
def test_scenarios_bit_exact(golden_dir, tag, n, rule):
    g = np.load(os.path.join(golden_dir, "scenarios.npz"))
    for phase in ("test", "val", "train"):
        ref = g["%s_%s" % (tag, phase)]
        for case in range(ref.shape[0]):
            rb, hs = scenarios.generate_case(phase, case, n, rule)
            assert np.array_equal(hs[:, [0, 1, 5, 6]], ref[case])
            assert np.array_equal(rb, [0, -4, 0, 0, 0.3, 0, 4, 1.0, np.pi / 2])


def test_case_seeds():
    assert scenarios.case_seed("val", 3) == 3
    assert scenarios.case_seed("test", 0) == 1000
    assert scenarios.case_seed("train", 0) == 2000


def make_env(n=5, sim="circle_crossing"):
    cfg = env_config(sim__human_num=n)
    env = make("CrowdSim-v0")
    env.configure(cfg)
    env.test_sim = sim
    robot = Robot(cfg, "robot")
    torch.manual_seed(0)
    policy = policy_factory["actenvcarl"]()
    policy.configure(policy_config())
    robot.set_policy(policy)
    return env, robot, policy


def test_crowdsim_reset_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "scenarios.npz"))
    env, robot, policy = make_env()
    with pytest.raises(AttributeError, match="robot has to be set"):
        env.reset("test")
    env.set_robot(robot)
    with pytest.raises(AssertionError):
        env.reset("bogus")
    for case in range(6):
        ob = env.reset("test", case)
        got = np.array([[h.px, h.py, h.gx, h.gy] for h in env.humans])
        assert np.array_equal(got, g["circle5_test"][case])
        assert len(ob) == 5 and isinstance(ob[0], ObservableState)
        assert env.global_time == 0 and policy.time_step == 0.2 and robot.time_step == 0.2
    assert env.case_counter["test"] == 6                      # counter advanced and wraps mod case_size
    assert env.case_size == {"train": np.iinfo(np.uint32).max - 2000, "val": 100, "test": 500}
    assert (robot.px, robot.py, robot.gx, robot.gy, robot.theta) == (0, -4.0, 0, 4.0, np.pi / 2)
    env.reset("test", 499)
    assert env.case_counter["test"] == 0


def test_value_types():
    fs = FullState(1, 2, 3, 4, 0.3, 5, 6, 1.0, 0.5)
    os_ = ObservableState(7, 8, 9, 10, 0.4)
    assert fs + os_ == (1, 2, 3, 4, 0.3, 5, 6, 1.0, 0.5, 7, 8, 9, 10, 0.4)      # state.py:17-18,36-37
    JointState(fs, [os_])
    with pytest.raises(AssertionError):
        JointState(os_, [os_])
    assert ActionXY(1, 2).vx == 1


def test_policy_surface_and_errors():
    env, robot, policy = make_env()
    assert policy.name == "ACTENVCARL" and policy.trainable and policy.multiagent_training
    assert policy.gamma == 0.9 and policy.kinematics == "holonomic" and policy.query_env
    state = JointState(FullState(0, -4, 0, 0, 0.3, 0, 4, 1.0, np.pi / 2), [ObservableState(1, 1, 0, 0, 0.3)])
    with pytest.raises(AttributeError, match="Phase, device"):
        policy.predict(state)
    policy.phase, policy.device = "train", torch.device("cuda:0")
    with pytest.raises(AttributeError, match="Epsilon"):
        policy.predict(state)
    # goal short-circuit comes before anything else (multi_human_rl.py:321-322)
    policy.phase = "test"
    at_goal = JointState(FullState(0, 4, 0, 0, 0.3, 0, 4, 1.0, 0), [ObservableState(1, 1, 0, 0, 0.3)])
    assert policy.predict(at_goal) == ActionXY(0, 0)
    policy.build_action_space(1.0)
    assert len(policy.action_space) == 81 and policy.action_space[0] == ActionXY(0, 0)
    assert abs(policy.action_space[6].vx - 0.11904303084504313) < 1e-16
    with pytest.raises(RuntimeError, match="CUDA"):
        policy.set_device(torch.device("cpu"))


def test_state_dict_surface_matches_reference_checkpoints():
    _, _, policy = make_env()
    sd = policy.get_model().state_dict()
    assert sum(v.numel() for v in sd.values()) == 139352
    keys = list(sd.keys())
    for k, shape in wts.WEIGHT_SPEC + wts.DEAD_TEMPLATE_SPEC:
        assert k in keys and tuple(sd[k].shape) == tuple(shape)
    flat = policy.get_model().flat_weights()
    assert flat.shape == (wts.NUM_WEIGHTS,)
    rt = wts.unflatten(flat)
    assert all(np.array_equal(rt[k], sd[k].numpy()) for k, _ in wts.WEIGHT_SPEC)
    m2 = policy_factory["actenvcarl"]()
    m2.configure(policy_config())
    m2.get_model().load_state_dict(sd)                                            # checkpoint round trip
    assert np.array_equal(m2.get_model().flat_weights(), flat)


def test_module_forward_batch_global_halting_matches_oracle_per_sample_at_batch_one(oracle):
    _, _, policy = make_env()
    m = policy.get_model()
    W = m.flat_weights()
    rng = np.random.RandomState(0)
    for _ in range(5):
        st = rng.randn(5, 13).astype(np.float32)
        st[:, 1:6] = st[0, 1:6]
        with torch.no_grad():
            v, _ = m.forward_torch(torch.from_numpy(st).unsqueeze(0))
        vo, ko = oracle.value_forward(W, st)
        assert abs(v.item() - vo) <= 2e-6 * max(abs(vo), 1e-2) + 2e-7
        assert m.get_step_cnt() == ko


def test_replay_memory_ring():
    mem = ReplayMemory(3)
    for i in range(5):
        mem.push(i)
    assert len(mem) == 3 and mem.is_full() and sorted(mem.memory) == [2, 3, 4] and mem[mem.position - 1] == 4


def test_default_configs_have_reference_keys():
    e, p, t = env_config(readme_reward_flags=False), policy_config(), train_config()
    assert e.getfloat("reward", "human_timestep") == 0.6 and e.getint("env", "test_size") == 500
    assert env_config().getfloat("reward", "human_timestep") == 0.5
    assert p.get("actenvcarl", "action_dims") == "100, 50, 1" and p.get("actenvcarl", "act_fixed") is False
    assert t.getint("trainer", "batch_size") == 100 and t.getint("train", "capacity") == 100000


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from aemcarl_b200._lib import AemError
    from aemcarl_b200.engine import Engine
    with pytest.raises(AemError, match="no CPU fallback"):
        Engine(0)


def test_device_replay_memory_ring_matches_reference_ring():
    """DeviceReplayMemory (utils/vec_explorer.py) keeps the ring semantics of memory.py:4-28, also for batched pushes
    that wrap around or exceed the capacity"""
    import torch
    from aemcarl_b200.crowd_nav.utils.memory import ReplayMemory
    from aemcarl_b200.crowd_nav.utils.vec_explorer import DeviceReplayMemory
    cap, n = 7, 2
    ref, dev = ReplayMemory(cap), DeviceReplayMemory(cap, n, "cpu")
    counter = 0
    for m in (3, 2, 4, 1, 9, 7, 16, 5):
        st = torch.arange(counter, counter + m, dtype=torch.float32).reshape(m, 1, 1).expand(m, n, 13).contiguous()
        va = torch.arange(counter, counter + m, dtype=torch.float32).reshape(m, 1) * 0.5
        counter += m
        for i in range(m):
            ref.push((st[i], va[i]))
        dev.push_batch(st, va)
        assert len(dev) == len(ref) and dev.position == ref.position and dev.is_full() == ref.is_full()
        for i in range(len(ref)):
            assert torch.equal(dev[i][0], ref[i][0]) and torch.equal(dev[i][1], ref[i][1])
    g = torch.Generator().manual_seed(0)
    xs, vs = dev.sample(4, g)
    assert xs.shape == (4, n, 13) and vs.shape == (4, 1) and len(set(vs.reshape(-1).tolist())) == 4
    seen = torch.cat([v for _, v in dev.epoch_batches(3, drop_last=True, generator=g)])
    assert seen.numel() == 6 and len(set(seen.reshape(-1).tolist())) == 6
    dev.clear()
    assert len(dev) == 0


def test_control_flow_free_act_equals_the_reference_loop():
    """ATCBasic.forward_static (CUDA-graph capturable) == ATCBasic.forward (components.py:107-137, batch-global halting)
    for weights that halt after 1, 2-3 and 3 steps, values and gradients"""
    import torch
    from aemcarl_b200 import weights as wts
    from aemcarl_b200.crowd_nav.default_configs import policy_config
    from aemcarl_b200.crowd_nav.policy.policy_factory import policy_factory
    policy = policy_factory["actenvcarl"]()
    policy.configure(policy_config())
    model = policy.get_model()
    counts = set()
    for seed, pb in ((1, None), (2, 4.0), (3, -1.0), (4, -0.13), (4, 0.9)):
        model.load_flat_weights(wts.synthetic_weights(seed, pb))
        x = torch.randn(9, 5, 13, generator=torch.Generator().manual_seed(seed))
        grads = []
        for static in (False, True):
            model.zero_grad()
            v, _ = model.forward_torch(x, static=static)
            counts.add(int(model.step_cnt))
            v.sum().backward()
            grads.append((v.detach().clone(), [p.grad.clone() for p in model.parameters() if p.grad is not None]))
        assert torch.equal(grads[0][0], grads[1][0])
        assert all(torch.allclose(a, b, rtol=1e-6, atol=1e-8) for a, b in zip(grads[0][1], grads[1][1]))
    assert len(counts) >= 2


def test_summary_lines_are_parsed_by_the_reference_plot_tool():
    """output.log compatibility: the TRAIN / VAL summary lines match what crowd_nav/utils/plot.py:38-55
    extracts (episode, success rate, collision rate, nav time, total reward)"""
    import re
    from aemcarl_b200.crowd_nav.utils.explorer import summary_line
    fields = (r" in episode (?P<episode>\d+) has success rate: (?P<sr>[0-1].\d+), collision rate: (?P<cr>[0-1].\d+), "
              r"nav time: (?P<time>\d+.\d+), total reward: (?P<reward>[-+]?\d+.\d+)")
    for phase, tag in (("train", "TRAIN"), ("val", "VAL  ")):
        line = summary_line(phase, 1200, 0.83, 0.05, 11.25, -0.03125)
        m = re.search(tag + fields, line)
        assert m, line
        assert (int(m["episode"]), float(m["sr"]), float(m["cr"]), float(m["time"])) == (1200, 0.83, 0.05, 11.25)
        assert float(m["reward"]) == -0.0312 or float(m["reward"]) == -0.0313
    assert summary_line("test", None, 1.0, 0.0, 10.5, 0.3).startswith("TEST  has success rate: 1.00, collision rate: 0.00")


def test_fused_training_step_fails_loudly_without_cuda():
    """Trainer(fused=True) is the CUDA kernel path (csrc/train.cu); on a CPU model it must raise, not fall back to autograd"""
    import torch
    from aemcarl_b200.crowd_nav.common.value_network import ValueNetworkTransformerAndGRU
    from aemcarl_b200.crowd_nav.utils.trainer import Trainer
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    model = ValueNetworkTransformerAndGRU().eval()
    tr = Trainer(model, None, torch.device("cpu"), 4, fused=True)
    tr.set_learning_rate(1e-3, "Adam")
    with pytest.raises(Exception) as e:
        tr._step(torch.zeros((4, 5, 13)), torch.zeros((4, 1)))
    assert "CUDA" in str(e.value) or "cuda" in str(e.value)
    model.train()                                   # dropout-active training is part of the fused kernel's contract now:
    with pytest.raises(RuntimeError, match="CUDA"):  # the same loud failure without a GPU, no silent autograd path
        tr._step(torch.zeros((4, 5, 13)), torch.zeros((4, 1)))
    fixed = ValueNetworkTransformerAndGRU(act_fixed=True, act_steps=2).eval()
    tr2 = Trainer(fixed, None, torch.device("cpu"), 4, fused=True)
    tr2.set_learning_rate(1e-3, "Adam")
    with pytest.raises(RuntimeError, match="CUDA"):  # act_fixed (<= 3 steps) is inside the fused kernels: same loud failure
        tr2._step(torch.zeros((4, 5, 13)), torch.zeros((4, 1)))
    fixed5 = ValueNetworkTransformerAndGRU(act_fixed=True, act_steps=5).eval()
    tr3 = Trainer(fixed5, None, torch.device("cpu"), 4, fused=True)
    tr3.set_learning_rate(1e-3, "Adam")
    with pytest.raises(NotImplementedError):         # more steps than the kernels' stash holds
        tr3._step(torch.zeros((4, 5, 13)), torch.zeros((4, 1)))


def test_device_scenario_params_struct_matches_header():
    import ctypes
    from aemcarl_b200._lib import ScenarioParams
    assert ctypes.sizeof(ScenarioParams) == 7 * 8 + 8


def test_per_sample_halting_forward_equals_one_state_at_a_time():
    """forward_torch(per_sample=True) on a batch == the module called on every state alone (the reference's batch-1 calls at
    explorer.py:137 / multi_human_rl.py:351): the rule VecExplorer's train()-mode target network uses"""
    import torch
    from aemcarl_b200 import weights as wts
    from aemcarl_b200.crowd_nav.common.value_network import ValueNetworkTransformerAndGRU
    rng = np.random.RandomState(5)
    states = rng.uniform(-1, 1, size=(40, 5, 13))
    states[:, :, :6] = states[:, :1, :6]
    st = torch.as_tensor(states, dtype=torch.float64)
    model = ValueNetworkTransformerAndGRU().double().eval()
    model.load_flat_weights(wts.synthetic_weights(3, -0.05))          # mixed halting: 2 and 3 steps
    with torch.no_grad():
        batch, _ = model.forward_torch(st, per_sample=True)
        steps = model.step_cnt.clone()
        single, ks = [], []
        for i in range(st.shape[0]):
            v, _ = model.forward_torch(st[i:i + 1])
            single.append(float(v))
            ks.append(int(model.step_cnt))
    assert len(set(ks)) >= 2 and steps.tolist() == ks
    assert np.abs(batch.numpy().reshape(-1) - np.array(single)).max() <= 1e-12
    model.train()                                                     # dropout live: a different, finite answer every call
    with torch.no_grad():
        a, _ = model.forward_torch(st, per_sample=True)
        b, _ = model.forward_torch(st, per_sample=True)
    assert torch.isfinite(a).all() and not torch.equal(a, b)
