"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) under oracle/ref_shim.py.

Run in the build container only:   python -m oracle.make_golden [--quick]
The fixtures are committed; the GPU box never needs /root/reference.

Documented deviations of the reference run (SURVEY.md hazards 1,2,3):
  * model.eval()  (the reference never calls it; dropout p=0.1 would make values random)
  * rvo2 is our own FP32 C restatement (oracle/aem_oracle.c) behind a PyRVOSimulator-shaped stub, so ORCA
    outputs in these fixtures are self-consistent, NOT pinned against the real Python-RVO2.
Weights are not stored: they are re-created from aemcarl_b200.weights.synthetic_weights(seed, ponder_bias).
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from aemcarl_b200 import weights as wts  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (name, seed, ponder_bias): steer ACT halting counts to 1 / 2 / 3 / mixed
WEIGHT_SETS = [("k2", 1, None), ("k1", 2, 4.0), ("k3", 3, -1.0), ("mixed", 4, -0.13)]


def load_weights(policy, flat):
    import torch
    sd = policy.get_model().state_dict()
    new = {k: torch.from_numpy(v.copy()) for k, v in wts.unflatten(flat).items()}
    for k, s in wts.DEAD_TEMPLATE_SPEC:
        new[k] = sd[k]
    policy.get_model().load_state_dict(new)


def full_states(env):
    r = env.robot
    robot = np.array([r.px, r.py, r.vx, r.vy, r.radius, r.gx, r.gy, r.v_pref, r.theta], dtype=np.float64)
    humans = np.array([[h.px, h.py, h.vx, h.vy, h.radius, h.gx, h.gy, h.v_pref] for h in env.humans],
                      dtype=np.float64)
    return robot, humans


class Recorder:
    """forward hook: raw network value + ACT step count of every per-action forward of predict()"""

    def __init__(self, model):
        self.vals, self.steps = [], []
        model.register_forward_hook(self)
        self.model = model

    def __call__(self, module, inp, out):
        self.vals.append(float(out[0].data.item()))
        self.steps.append(int(module.get_step_cnt()))

    def pop(self):
        v, s = np.array(self.vals, dtype=np.float32), np.array(self.steps, dtype=np.uint8)
        self.vals, self.steps = [], []
        return v, s


INFO_CODE = {"Nothing": 0, "Danger": 1, "ReachGoal": 2, "Collision": 3, "Timeout": 4}


def decision_record(env, policy, rec):
    """Everything predict() sees for the current env state: per-action lookahead of the reference env."""
    robot, humans = full_states(env)
    rewards, infos, dmins, hnext = [], [], [], None
    for a in policy.action_space:
        ob, reward, done, info = env.onestep_lookahead(a, True)
        rewards.append(float(reward))
        infos.append(INFO_CODE[type(info).__name__])
        dmins.append(float(getattr(info, "min_dist", np.inf)))
        if hnext is None:
            hnext = np.array([[o.px, o.py, o.vx, o.vy, o.radius] for o in ob], dtype=np.float64)
    return dict(robot=robot, humans=humans, global_time=float(env.global_time), rewards=np.array(rewards),
                infos=np.array(infos, dtype=np.int32), danger_dmin=np.array(dmins), humans_next=hnext)


def run_episode(env, robot, policy, rec, case, max_steps, record_every):
    """Explorer.run_k_episodes inner loop (explorer.py:46-56) with recording."""
    from crowd_sim.envs.utils.state import JointState
    ob = env.reset("test", case)
    if policy.action_space is None:
        policy.build_action_space(robot.v_pref)
    done, step = False, 0
    records, actions, rewards = [], [], []
    info = None
    while not done and step < max_steps:
        want = (step % record_every == 0)
        if want:
            r = decision_record(env, policy, rec)
        rec.pop()
        action = robot.act(ob)
        net_v, net_s = rec.pop()
        idx = -1
        for i, a in enumerate(policy.action_space):
            if a is action:
                idx = i
        if want and idx >= 0:
            r.update(values=np.array(policy.action_values, dtype=np.float64), net_values=net_v, act_steps=net_s,
                     best=idx, case=case, step=step)
            records.append(r)
        ob, reward, done, info = env.step(action)
        actions.append(idx)
        rewards.append(float(reward))
        step += 1
    outcome = INFO_CODE[type(info).__name__] if done else -1
    return records, np.array(actions, dtype=np.int32), np.array(rewards), outcome, float(env.global_time)


def stack(records, keys):
    return {k: np.stack([r[k] for r in records]) for k in keys}


def gen_decisions(tag, human_num, sim, cases, wname, max_steps, record_every, visible=False):
    t0 = time.time()
    _, seed, pb = [w for w in WEIGHT_SETS if w[0] == wname][0]
    policy = ref_shim.make_policy(seed=0)
    load_weights(policy, wts.synthetic_weights(seed, pb))
    env, robot, _ = ref_shim.make_env(human_num=human_num, test_sim=sim, visible=visible)
    robot.set_policy(policy)
    policy.set_env(env)
    rec = Recorder(policy.get_model())
    allrec, episodes = [], {}
    for case in cases:
        records, actions, rewards, outcome, gtime = run_episode(env, robot, policy, rec, case, max_steps,
                                                                record_every)
        allrec += records
        episodes["actions_case%d" % case] = actions
        episodes["rewards_case%d" % case] = rewards
        episodes["outcome_case%d" % case] = np.array([outcome, len(actions)], dtype=np.int32)
        episodes["time_case%d" % case] = np.array(gtime)
        print("  %s case %d: %d steps outcome %d (%.1fs)" % (tag, case, len(actions), outcome, time.time() - t0))
    keys = ["robot", "humans", "global_time", "rewards", "infos", "danger_dmin", "humans_next", "values",
            "net_values", "act_steps", "best", "case", "step"]
    data = stack(allrec, keys)
    data.update(episodes)
    data["weight_seed"] = np.array(seed)
    data["ponder_bias"] = np.array(np.nan if pb is None else pb)
    data["robot_visible"] = np.array(int(visible))
    data["cases"] = np.array(cases, dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "decisions_%s.npz" % tag), **data)
    return data


def gen_values_other_weights(base, tag):
    """Re-evaluate the recorded lookahead states with the other weight sets (ACT counts 1 / 3 / mixed)."""
    import torch
    policy = ref_shim.make_policy(seed=0)
    model = policy.get_model()
    out = {}
    for wname, seed, pb in WEIGHT_SETS:
        load_weights(policy, wts.synthetic_weights(seed, pb))
        policy.build_action_space(1.0)
        policy.time_step = 0.2
        from crowd_sim.envs.utils.state import FullState, ObservableState
        R = base["robot"].shape[0]
        vals = np.zeros((R, 81), dtype=np.float32)
        steps = np.zeros((R, 81), dtype=np.uint8)
        for r in range(min(R, 6)):
            rb = base["robot"][r]
            self_state = FullState(*[float(x) for x in rb])
            hn = [ObservableState(*[float(x) for x in h]) for h in base["humans_next"][r]]
            for ai, action in enumerate(policy.action_space):
                nxt = policy.propagate(self_state, action)
                batch = torch.cat([torch.Tensor([nxt + h]) for h in hn], dim=0)   # multi_human_rl.py:351-355
                with torch.no_grad():
                    v, _ = model(policy.rotate(batch).unsqueeze(0))
                vals[r, ai] = v.item()
                steps[r, ai] = model.get_step_cnt()
        out["net_values_" + wname] = vals[:min(R, 6)]
        out["act_steps_" + wname] = steps[:min(R, 6)]
        out["seed_" + wname] = np.array(seed)
        out["ponder_bias_" + wname] = np.array(np.nan if pb is None else pb)
        print("  %s weights %s: steps histogram %s" % (tag, wname, np.bincount(steps[:min(R, 6)].ravel(), minlength=4)))
    out["robot"] = base["robot"][:6]
    out["humans_next"] = base["humans_next"][:6]
    np.savez_compressed(os.path.join(OUT, "values_%s.npz" % tag), **out)


def gen_update_memory():
    """Explorer.update_memory (crowd_nav/utils/explorer.py:114-151) of the UNMODIFIED reference on one recorded episode: the
    imitation-learning targets (discounted Monte-Carlo return, :126-130) and the RL targets (r + gamma_bar V_target(s'),
    :132-139), with the (rotated state, value) pairs the reference pushes into its ReplayMemory."""
    import torch
    _, seed, pb = WEIGHT_SETS[0]
    policy = ref_shim.make_policy(seed=0)               # installs the shim: the reference packages import from here on
    from crowd_nav.utils.explorer import Explorer
    from crowd_nav.utils.memory import ReplayMemory
    from crowd_sim.envs.utils.state import JointState
    load_weights(policy, wts.synthetic_weights(seed, pb))
    env, robot, _ = ref_shim.make_env(human_num=5)
    robot.set_policy(policy)
    policy.set_env(env)
    ob = env.reset("test", 3)
    if policy.action_space is None:
        policy.build_action_space(robot.v_pref)
    joint, rewards = [], []
    for step in range(14):
        state = JointState(robot.get_full_state(), ob)
        joint.append(state)
        action = robot.act(ob)
        ob, reward, done, info = env.step(action)
        rewards.append(float(reward))
        if done:
            break
    # the environment's rewards of a quiet episode are all zero: the targets are checked on a synthetic reward sequence
    rewards = [float(np.round(0.1 * np.sin(1.7 * i), 3)) for i in range(len(joint))]
    rewards[-1] = 1.0
    rb = np.array([[s.self_state.px, s.self_state.py, s.self_state.vx, s.self_state.vy, s.self_state.radius, s.self_state.gx,
                    s.self_state.gy, s.self_state.v_pref, s.self_state.theta] for s in joint])
    hs = np.array([[[h.px, h.py, h.vx, h.vy, h.radius] for h in s.human_states] for s in joint])
    out = dict(robot=rb, humans_obs=hs, rewards=np.array(rewards), weight_seed=np.array(seed), gamma=np.array(policy.gamma),
               time_step=np.array(robot.time_step), v_pref=np.array(robot.v_pref))
    for il in (True, False):
        memory = ReplayMemory(1000)
        ex = Explorer(env, robot, torch.device("cpu"), memory, policy.gamma, target_policy=policy)
        ex.update_target_model(policy.get_model())
        ex.target_model.eval()
        states = joint if il else [policy.transform(s) for s in joint]
        ex.update_memory(states, [None] * len(states), rewards, imitation_learning=il)
        tag = "il" if il else "rl"
        out[tag + "_states"] = np.stack([m[0].numpy() for m in memory.memory])
        out[tag + "_values"] = np.array([float(m[1]) for m in memory.memory])
    # Trainer (crowd_nav/utils/trainer.py:22-86) of the reference on the RL pairs above: one optimize_batch step whose batch is
    # the whole memory (so the DataLoader's shuffle only permutes it), for the three optimizers it offers
    from crowd_nav.utils.trainer import Trainer
    import copy
    for opt in ("Adam", "SGD", "RMSprop"):
        model = copy.deepcopy(policy.get_model())
        model.eval()
        tr = Trainer(model, memory, torch.device("cpu"), len(memory.memory))
        tr.set_learning_rate(0.001, opt)
        torch.manual_seed(0)
        loss = tr.optimize_batch(1)
        out["trainer_%s_loss" % opt] = np.array(float(loss))
        out["trainer_%s_params" % opt] = wts.flatten_state_dict(model.state_dict())[::8].copy()    # every 8th parameter: small fixture
    np.savez_compressed(os.path.join(OUT, "update_memory.npz"), **out)
    print("  update_memory: %d steps, IL values %.4f .. %.4f, RL values %.4f .. %.4f" % (
        len(rewards), out["il_values"].min(), out["il_values"].max(), out["rl_values"].min(), out["rl_values"].max()))
    return out


def gen_action_rotate():
    policy = ref_shim.make_policy(seed=0)
    policy.build_action_space(1.0)
    import torch
    acts = np.array([[a.vx, a.vy] for a in policy.action_space], dtype=np.float64)
    rng = np.random.RandomState(7)
    x = rng.uniform(-5, 5, size=(256, 14))
    x[:, 4] = rng.uniform(0.2, 0.5, 256)
    x[:, 13] = rng.uniform(0.2, 0.5, 256)
    x[:, 7] = rng.uniform(0.5, 1.5, 256)
    x32 = x.astype(np.float32)
    rot = policy.rotate(torch.from_numpy(x32)).numpy()
    np.savez_compressed(os.path.join(OUT, "action_rotate.npz"), actions=acts, speeds=np.array(policy.speeds),
                        rotations=np.array(policy.rotations), gamma_bar=np.array(pow(0.9, 0.2 * 1.0)),
                        rotate_in=x32, rotate_out=rot)


def gen_scenarios():
    out = {}
    for tag, n, sim, ncase in (("circle5", 5, "circle_crossing", 64), ("square10", 10, "square_crossing", 32),
                               ("circle18", 18, "circle_crossing", 16)):
        env, robot, _ = ref_shim.make_env(human_num=n, test_sim=sim)
        policy = ref_shim.make_policy(seed=0)
        robot.set_policy(policy)
        for phase in ("test", "val", "train"):
            res = []
            for case in range(ncase if phase == "test" else 4):
                env.reset(phase, case)
                res.append([[h.px, h.py, h.gx, h.gy] for h in env.humans])
            out["%s_%s" % (tag, phase)] = np.array(res, dtype=np.float64)
    # the 'mixed' rule (variable human count per case) and randomize_attributes: rows padded to 6 humans,
    # columns px, py, gx, gy, radius, v_pref; *_count = len(env.humans), *_human_num = env.human_num after reset
    for tag, sim, rand, ncase in (("mixed5", "mixed", False, 96), ("mixed5_rand", "mixed", True, 48),
                                  ("circle5_rand", "circle_crossing", True, 32), ("square5_rand", "square_crossing", True, 32)):
        env, robot, _ = ref_shim.make_env(human_num=5, test_sim=sim)
        env.randomize_attributes = rand
        policy = ref_shim.make_policy(seed=0)
        robot.set_policy(policy)
        rows = np.zeros((ncase, 6, 6), dtype=np.float64)
        count = np.zeros(ncase, dtype=np.int64)
        hnum = np.zeros(ncase, dtype=np.int64)
        for case in range(ncase):
            env.human_num = 5
            env.reset("test", case)
            count[case], hnum[case] = len(env.humans), env.human_num
            for i, h in enumerate(env.humans):
                rows[case, i] = [h.px, h.py, h.gx, h.gy, h.radius, h.v_pref]
        out[tag + "_rows"], out[tag + "_count"], out[tag + "_human_num"] = rows, count, hnum
    np.savez_compressed(os.path.join(OUT, "scenarios.npz"), **out)


def gen_reward():
    """gridmap.build_map_cpu + compute_occupied_probability (crowd_sim.py:159-219) on random layouts."""
    ref_shim.install()
    from crowd_sim.envs.crowd_sim import gridmap
    from crowd_sim.envs.utils.state import FullState, ObservableState
    policy = ref_shim.make_policy(seed=0)
    policy.build_action_space(1.0)
    acts = np.array([[a.vx, a.vy] for a in policy.action_space])
    rng = np.random.RandomState(11)
    out = {}
    for tag, (ats, hts, pv, dv) in (("readme", (0.4, 0.5, 2.0, 2.0)), ("config", (0.4, 0.6, 4.0, 3.0))):
        gm = gridmap(xlimit=12, ylimit=12, resolution=0.02, time_step=hts, position_variance=pv,
                     direction_variance=dv)
        robots, humans, probs = [], [], []
        for trial in range(24):
            n = [5, 10, 18][trial % 3]
            rp = rng.uniform(-3.5, 3.5, 2)
            hs = []
            for i in range(n):
                if i < 3:   # near the robot so the discs overlap blobs
                    p = rp + rng.uniform(-1.2, 1.2, 2)
                else:
                    p = rng.uniform(-4.5, 4.5, 2)
                ang, sp = rng.uniform(0, 2 * np.pi), rng.uniform(0, 1.2)
                if i == 1:
                    sp = 0.0     # stationary human -> empty blob
                hs.append([p[0], p[1], sp * np.cos(ang), sp * np.sin(ang), 0.3])
            if trial == 0:
                rp = np.array([-5.9, -5.8])   # lower clamp of the disc bounds
                hs[0][:2] = [-5.5, -5.6]
            hstates = [ObservableState(*h) for h in hs]
            gm.build_map_cpu(hstates, True)
            pr = []
            for a in acts:
                fs = FullState(rp[0] + a[0] * ats, rp[1] + a[1] * ats, 0.0, 0.0, 0.3, 0.0, 4.0, 1.0, 0.0)
                pr.append(gm.compute_occupied_probability(fs))
            robots.append(rp)
            humans.append(np.array(hs + [[0] * 5] * (18 - n)))
            probs.append(pr)
            out["%s_n" % tag] = np.array([[5, 10, 18][t % 3] for t in range(len(robots))], dtype=np.int32)
        out["%s_robot" % tag] = np.array(robots)
        out["%s_humans" % tag] = np.array(humans)
        out["%s_prob" % tag] = np.array(probs)
        out["%s_params" % tag] = np.array([ats, hts, pv, dv])
    np.savez_compressed(os.path.join(OUT, "reward_map.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    only = set(args.only.split(",")) if args.only else None

    def want(x):
        return only is None or x in only

    t0 = time.time()
    if want("basic"):
        gen_action_rotate()
        gen_scenarios()
        print("basic done %.1fs" % (time.time() - t0))
    if want("reward"):
        gen_reward()
        print("reward done %.1fs" % (time.time() - t0))
    if want("memory"):
        gen_update_memory()
        print("update_memory done %.1fs" % (time.time() - t0))
    if want("n5"):
        d5 = gen_decisions("circle5", 5, "circle_crossing", [0, 1, 2, 3], "k2", 200, 4)
        gen_values_other_weights(d5, "circle5")
    if want("n5mixed"):
        gen_decisions("circle5_mixed", 5, "circle_crossing", [7, 8], "mixed", 200, 6)
    if want("n5vis"):
        gen_decisions("circle5_visible", 5, "circle_crossing", [4], "k2", 40, 5, visible=True)
    if want("n10"):
        d10 = gen_decisions("square10", 10, "square_crossing", [0, 1], "k2", 36, 6)
        gen_values_other_weights(d10, "square10")
    if want("n18"):
        d18 = gen_decisions("circle18", 18, "circle_crossing", [0], "k2", 24, 6)
        gen_values_other_weights(d18, "circle18")
    print("all done %.1fs" % (time.time() - t0))


if __name__ == "__main__":
    main()
