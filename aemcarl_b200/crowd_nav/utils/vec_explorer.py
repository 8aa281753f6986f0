"""Vectorised Explorer with an on-device replay memory.

The reference's Explorer (crowd_nav/utils/explorer.py:22-151) plays one episode at a time and, for every
success / collision episode, pushes one (rotated state, target value) pair per step into a Python list.  Here B
episodes run in lock-step on a VecCrowdSim (one fused lookahead + step launch sequence per step for all of them), the
trajectories stay in HBM, the target values of all finished episodes are computed by ONE batched aem_value_forward
on the target network, and the pairs land in a ring buffer that is itself a pair of device tensors.

Semantics kept from the reference:
  * a step stores last_state = transform(state BEFORE the action) (multi_human_rl.py:385-386, explorer.py:53);
  * only episodes that end in ReachGoal or Collision are stored (explorer.py:78-80);
  * RL target (explorer.py:132-139): value_i = r_i + gamma^(dt * v_pref) * V_target(state_{i+1}), last step value = r;
  * IL target (explorer.py:126-130): value_i = sum_{s >= i} gamma^((s - i) * dt * v_pref) * r_s;
  * the statistics and the log line of run_k_episodes (explorer.py:82-108).
Differences (documented, unavoidable in a vectorised run): the pairs of a wave are pushed in case order after the wave
(the reference pushes episode by episode as they finish), so the ORDER of the pairs in the ring differs (the multiset per
episode is identical -- tests/test_gpu_explorer.py); epsilon-greedy draws come from a seeded torch.Generator on the device
instead of the global numpy stream; training scenario pools are drawn on the device (aem_generate_scenarios: the same seeded
MT19937 stream, circle-rule coordinates within an ulp of numpy's cos / sin), val / test pools on the host (bit-exact).
The loop itself never synchronises with the host except for one "anyone still playing?" flag every fourth step.
"""
import logging

import numpy as np
import torch

from aemcarl_b200.crowd_nav.utils.explorer import average, summary_line

INFO_DANGER, INFO_REACHGOAL, INFO_COLLISION, INFO_TIMEOUT = 1, 2, 3, 4


class DeviceReplayMemory(object):
    """Replay ring buffer (crowd_nav/utils/memory.py:4-28) held as two tensors on `device`:
    states [capacity, n, 13] f32 and values [capacity, 1] f32.  Works on any torch device (the ring logic is tested
    on the CPU); push() keeps the reference's overwrite-oldest order."""

    def __init__(self, capacity, human_num, device):
        self.capacity = int(capacity)
        self.device = torch.device(device)
        self.states = torch.zeros((self.capacity, human_num, 13), dtype=torch.float32, device=self.device)
        self.values = torch.zeros((self.capacity, 1), dtype=torch.float32, device=self.device)
        self.position = 0
        self.size = 0
        self.sample_seed = 0
        self._samples_drawn = 0

    def push(self, item):
        state, value = item
        self.push_batch(state.unsqueeze(0), value.reshape(1, 1))

    def push_batch(self, states, values):
        """append m pairs in order; like m successive push() calls"""
        m = int(states.shape[0])
        if m == 0:
            return
        states = states.to(self.device, torch.float32)
        values = values.to(self.device, torch.float32).reshape(m, 1)
        if m >= self.capacity:                      # only the last `capacity` survive
            first_kept = m - self.capacity
            end = (self.position + m) % self.capacity
            idx = (torch.arange(self.capacity, device=self.device) + end) % self.capacity
            self.states[idx] = states[first_kept:]
            self.values[idx] = values[first_kept:]
        else:
            idx = (torch.arange(m, device=self.device) + self.position) % self.capacity
            self.states[idx] = states
            self.values[idx] = values
        self.position = (self.position + m) % self.capacity
        self.size = min(self.capacity, self.size + m)

    def is_full(self):
        return self.size == self.capacity

    def __len__(self):
        return self.size

    def __getitem__(self, item):
        if item >= self.size:
            raise IndexError(item)
        return self.states[item], self.values[item]

    def clear(self):
        self.position = 0
        self.size = 0

    def sample(self, batch_size, generator=None):
        """one shuffled mini-batch without replacement (what next(iter(DataLoader(shuffle=True))) yields)"""
        k = min(int(batch_size), self.size)
        if self.device.type == "cuda" and generator is None:
            # aem_sample_batch: Floyd's subset draw + gather in one launch (randperm of the whole ring costs 0.1 ms per step)
            from aemcarl_b200.engine import get_engine
            self._samples_drawn += 1
            return get_engine(self.device.index or 0, role="trainer").sample_batch(self.states, self.values, self.size, k,
                                                                                   self.sample_seed, self._samples_drawn)
        perm = torch.randperm(self.size, device=self.device, generator=generator)[:k]
        return self.states[perm], self.values[perm]

    def epoch_batches(self, batch_size, drop_last=True, generator=None):
        perm = torch.randperm(self.size, device=self.device, generator=generator)
        stop = self.size - (self.size % batch_size if drop_last else 0)
        for i in range(0, stop, batch_size):
            idx = perm[i:i + batch_size]
            yield self.states[idx], self.values[idx]


class VecExplorer(object):
    """run_k_episodes for a batch of environments.  `policy` is the ACTENVCARL policy (its engine drives the lookahead),
    `num_envs` episodes are in flight at once."""

    def __init__(self, policy, num_envs, env_params, memory=None, gamma=None, human_num=5, rule="circle_crossing",
                 seed=0, scenario_kw=None, il_safety_space=0.0, device_scenarios=None, target_train_mode=False):
        self.policy = policy
        self.env_params = env_params          # aem_env_params of the configured CrowdSim (env.env_params())
        self.il_safety_space = float(il_safety_space)     # [imitation_learning] safety_space of the ORCA teacher
        self.B = int(num_envs)
        self.n = int(human_num)
        self.rule = rule
        self.memory = memory
        self.gamma = policy.gamma if gamma is None else gamma
        self.device = policy.device
        self.target_model = None
        # True: the target network is evaluated in train() mode, dropout live, like the reference's (it never calls eval(),
        # so the V_target of explorer.py:132-139 is a stochastic forward of one state at a time); default: the eval()-mode
        # network through aem_value_forward
        self.target_train_mode = bool(target_train_mode)
        self.scenario_kw = scenario_kw
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed))
        self.case_counter = {"train": 0, "val": 0, "test": 0}
        self.last_stats = None
        self._target_engine = None
        # scenario pools: None = drawn on the device for the training phase (aem_generate_scenarios; 4096 cases cost ~1 s in
        # the host generator), on the host (bit-exact with the reference) for val / test; True / False force one of them
        self.device_scenarios = device_scenarios
        # when the cases of a wave have run out, the kernels keep running on the environments still playing only, re-packed to
        # the front whenever fewer than this fraction of the current batch is left (None: never re-pack)
        self.compact_below = 0.75

    def update_target_model(self, target_model):
        import copy
        self.target_model = copy.deepcopy(target_model)

    def _target_values(self, states):
        """V_target of rotated states [m, n, 13] with one batched kernel launch (explorer.py:137, batch-1 there)"""
        from aemcarl_b200.engine import get_engine
        if self.target_model is None:
            raise ValueError("update_target_model() has not been called")
        if self.target_train_mode:
            # torch forward with live dropout; every state halts on its own as in the reference's batch-1 calls
            self.target_model.train()
            out = []
            with torch.no_grad():
                for i in range(0, states.shape[0], 8192):
                    out.append(self.target_model.forward_torch(states[i:i + 8192].contiguous(), per_sample=True)[0].reshape(-1))
            return torch.cat(out) if out else states.new_zeros((0,))
        if self._target_engine is None:
            self._target_engine = get_engine(self.device.index or 0, role="target")
        eng = self.target_model.sync_engine(self._target_engine)
        eng.set_precision(getattr(self.policy.engine(), "precision", 0))
        v, _ = eng.value_forward(states.contiguous())
        return v.reshape(-1)

    def run_k_episodes(self, k, phase, update_memory=False, imitation_learning=False, episode=None, epsilon=None,
                       first_case=None, print_failure=False):
        from aemcarl_b200.vec_env import VecCrowdSim
        if update_memory and (self.memory is None or self.gamma is None):
            raise ValueError("Memory or gamma value is not set!")
        pol = self.policy
        pol.set_phase(phase)
        eng = pol.engine()
        eng.set_env_params(self.env_params)
        first = self.case_counter[phase] if first_case is None else int(first_case)
        B = min(self.B, int(k))
        if pol.action_space is None:                 # before the env: its per-action buffers are sized by the action space
            pol.build_action_space(float((self.scenario_kw or {}).get("robot_v_pref", 1.0)))
            eng = pol.engine()
            eng.set_env_params(self.env_params)
        on_device = self.device_scenarios if self.device_scenarios is not None else phase == "train"
        env = VecCrowdSim(eng, B, human_num=self.n, rule=self.rule, phase=phase, first_case=first, pool_size=int(k),
                          gamma=self.gamma, scenario_kw=self.scenario_kw,
                          device_scenarios=on_device and self.rule in ("circle_crossing", "square_crossing"))
        dev = self.device
        dt = eng.params.time_step
        v_pref = float(env.pool_robot_h[0, 7])
        gbar = pow(self.gamma, dt * v_pref)
        tmax = int(round(eng.params.time_limit / dt)) + 2
        eps = float(epsilon if epsilon is not None else (pol.epsilon or 0.0)) if phase == "train" else 0.0

        # Everything below stays on the device and is issued without a host synchronisation: trajectories are written
        # straight into per-CASE rows (row k is a dummy for environments that ran out of cases), finished environments
        # take the next case in environment order (prefix sum over the done flags = the order of the serial loop
        # `for b in fin`), and the host looks at the device only every `check_every` steps to see whether any
        # environment is still playing.
        k = int(k)
        traj_s = torch.zeros((k + 1, tmax, self.n, 13), dtype=torch.float32, device=dev)
        traj_r = torch.zeros((k + 1, tmax), dtype=torch.float64, device=dev)
        traj_info = torch.zeros((k + 1, tmax), dtype=torch.int32, device=dev)
        traj_dmin = torch.zeros((k + 1, tmax), dtype=torch.float64, device=dev)
        ep_len = torch.zeros((k + 1,), dtype=torch.int64, device=dev)
        ep_code = torch.full((k + 1,), -1, dtype=torch.int32, device=dev)
        ep_time = torch.zeros((k + 1,), dtype=torch.float64, device=dev)
        step_idx = torch.zeros((B,), dtype=torch.int64, device=dev)
        case_of_env = torch.arange(B, device=dev, dtype=torch.int64)      # pool index each env is playing
        active = torch.ones((B,), dtype=torch.bool, device=dev)
        next_case = torch.full((), B, dtype=torch.int64, device=dev)
        env_steps_d = torch.zeros((), dtype=torch.int64, device=dev)
        dummy = torch.full((B,), k, dtype=torch.int64, device=dev)
        zero_t = torch.zeros((), dtype=torch.float64, device=dev)
        disc = torch.pow(torch.tensor(gbar, dtype=torch.float64, device=dev),
                         torch.arange(tmax, device=dev, dtype=torch.float64))
        slot = torch.arange(B, device=dev)                  # original environment index of every running environment
        check_every = 4
        it = 0

        il_eng = self._teacher_engine() if imitation_learning else None
        while True:
            cur = eng.transform(env.robot, env.humans)                      # last_state of every env
            if imitation_learning:
                reward, done, info, dmin = self._teacher_step(env, eng, il_eng, v_pref)
            else:
                best = env.lookahead()
                if eps > 0.0:
                    # drawn for the ORIGINAL batch and indexed by slot, so re-packing does not change any decision
                    explore = torch.rand((B,), device=dev, generator=self.gen)[slot] < eps
                    rnd = torch.randint(0, eng.A, (B,), device=dev, generator=self.gen, dtype=torch.int32)[slot]
                    # the goal short-circuit (action 0 inside the goal disc) precedes the epsilon draw (multi_human_rl.py:321)
                    at_goal = torch.linalg.norm(env.robot[:, 0:2] - env.robot[:, 5:7], dim=1) < env.robot[:, 4]
                    best = torch.where(explore & ~at_goal, rnd, best)
                reward, done, info, dmin = env.apply(best, auto_reset=False)
            row = torch.where(active, case_of_env, dummy)
            traj_s[row, step_idx] = cur
            traj_r[row, step_idx] = reward
            traj_info[row, step_idx] = info
            traj_dmin[row, step_idx] = dmin
            env_steps_d += active.sum()
            step_idx += 1
            step_idx.clamp_(max=tmax - 1)               # envs without a case left keep stepping harmlessly
            fin = active & (done != 0)
            erow = torch.where(fin, case_of_env, dummy)
            ep_len[erow] = step_idx
            ep_code[erow] = info
            ep_time[erow] = env.global_time
            newcase = next_case + torch.cumsum(fin, dim=0) - 1              # next cases in environment order
            next_case += fin.sum()
            has = fin & (newcase < k)
            case_of_env = torch.where(has, newcase, case_of_env)
            active = active & ~(fin & ~has)
            src = newcase.clamp(max=k - 1)
            env.robot.copy_(torch.where(has[:, None], env.pool_robot[src], env.robot))
            env.humans.copy_(torch.where(has[:, None, None], env.pool_humans[src], env.humans))
            env.global_time.copy_(torch.where(has, zero_t, env.global_time))
            step_idx = torch.where(fin, torch.zeros_like(step_idx), step_idx)
            it += 1
            if it % check_every == 0:                       # the only host synchronisation of the loop
                n_act = int(active.sum().item())
                if n_act == 0:
                    break
                if self.compact_below and n_act <= self.compact_below * env.B and env.B > 64:
                    # The cases have run out and part of the batch has finished for good (an environment only goes inactive
                    # when no case is left, so nothing is assigned any more): move the environments still playing to the front
                    # and let the kernels run on them alone.
                    keep = torch.nonzero(active).reshape(-1)
                    for t in (env.robot, env.humans, env.global_time):
                        t[:n_act] = t[keep]
                    step_idx, case_of_env, dummy, slot = step_idx[keep], case_of_env[keep], dummy[:n_act], slot[keep]
                    active = torch.ones((n_act,), dtype=torch.bool, device=dev)
                    env = env.narrow(n_act)

        # ---- one pass over the finished wave: statistics and the replay-memory pairs ----
        lens = ep_len[:k]
        live = torch.arange(tmax, device=dev).unsqueeze(0) < lens.unsqueeze(1)
        cum_reward = (traj_r[:k] * disc.unsqueeze(0) * live).sum(dim=1).cpu().numpy()
        codes = ep_code[:k].cpu().numpy().astype(np.int64)
        nav_time = ep_time[:k].cpu().numpy()
        danger = live & (traj_info[:k] == INFO_DANGER)
        too_close = int(danger.sum().item())
        min_dist = traj_dmin[:k][danger].cpu().numpy().tolist()
        env_steps = int(env_steps_d.item())
        pushed = 0
        store = np.nonzero(np.isin(codes, (INFO_REACHGOAL, INFO_COLLISION)))[0]
        if update_memory and len(store):
            pushed = self._store(traj_s, traj_r, ep_len, store, gbar, imitation_learning)

        self.case_counter[phase] = (first + k) % max(1, self._case_size(phase))
        success = int((codes == INFO_REACHGOAL).sum())
        collision = int((codes == INFO_COLLISION).sum())
        timeout = int((codes == INFO_TIMEOUT).sum())
        assert success + collision + timeout == k
        time_limit = eng.params.time_limit
        s_times = nav_time[codes == INFO_REACHGOAL]
        avg_nav_time = float(s_times.mean()) if len(s_times) else time_limit
        logging.info(summary_line(phase, episode, success / k, collision / k, avg_nav_time, average(list(cum_reward))))
        if phase in ["val", "test"]:
            total_time = (nav_time[codes != INFO_TIMEOUT].sum() + time_limit * timeout) * dt
            logging.info("Frequency of being in danger: %.2f and average min separate distance in danger: %.2f",
                         too_close / total_time, average(min_dist))
        if print_failure:
            logging.info("Collision cases: " + " ".join(str(x) for x in np.nonzero(codes == INFO_COLLISION)[0]))
            logging.info("Timeout cases: " + " ".join(str(x) for x in np.nonzero(codes == INFO_TIMEOUT)[0]))
        self.last_stats = dict(success=success, collision=collision, timeout=timeout, too_close=too_close,
                               avg_nav_time=avg_nav_time, cumulative_rewards=list(cum_reward), codes=codes,
                               env_steps=env_steps, pushed=pushed)
        return self.last_stats

    def _case_size(self, phase):
        return {"train": np.iinfo(np.uint32).max - 2000, "val": 100, "test": 500}[phase]

    def _teacher_engine(self):
        """a second engine whose ORCA parameters are the teacher's (orca.py:82-132 with [imitation_learning] safety_space)"""
        import copy
        from aemcarl_b200.engine import get_engine
        il = get_engine(self.device.index or 0, role="il_teacher")
        p = copy.copy(self.env_params)
        p.robot_visible = 0
        p.orca_safety_space = self.il_safety_space
        il.set_env_params(p)
        return il

    def _teacher_step(self, env, eng, il_eng, max_speed):
        """robot.act(ob) with the ORCA teacher + env.step(action) for every env (explorer.py:51-53 in the imitation phase):
        the humans move by their own ORCA solve (robot seen only if visible); the robot's action is agent 0 of ITS
        simulator = the robot appended to the humans as one more ORCA agent, solved with the teacher's safety space."""
        eng.orca(env.humans, env.robot, out=env.new_vel)
        agents = torch.cat([env.humans, env.robot[:, None, :8]], dim=1).contiguous()
        action = il_eng.orca(agents, env.robot)[:, self.n, :].contiguous()
        return eng.env_step_xy(env.robot, env.humans, env.new_vel, env.global_time, action, max_speed)

    def _store(self, traj_s, traj_r, step_idx, envs, gbar, imitation_learning=False):
        """targets of the finished episodes `envs` -> replay memory.  RL (explorer.py:132-139): one batched target
        forward; imitation (explorer.py:126-130): discounted Monte-Carlo return of the rest of the episode."""
        dev = self.device
        eb = torch.from_numpy(np.asarray(envs)).to(dev)
        lens = step_idx[eb]
        tmax = traj_s.shape[1]
        t = torch.arange(tmax, device=dev).unsqueeze(0)
        valid = t < lens.unsqueeze(1)                          # [m, tmax] steps of each episode
        has_next = t < (lens - 1).unsqueeze(1)
        states = traj_s[eb]                                    # [m, tmax, n, 13]
        rewards = traj_r[eb]
        values = rewards.clone()
        if imitation_learning:
            disc = torch.pow(torch.tensor(gbar, dtype=torch.float64, device=dev), t.to(torch.float64))     # gbar^s
            tail = torch.flip(torch.cumsum(torch.flip(rewards * disc * valid, dims=[1]), dim=1), dims=[1])
            values = tail / disc                                   # sum_{s >= i} gbar^(s - i) r_s
        elif has_next.any():
            nxt = torch.roll(states, -1, dims=1)[has_next]     # state_{i+1} of every non-final step
            v = self._target_values(nxt).to(torch.float64)
            values[has_next] = rewards[has_next] + gbar * v
        self.memory.push_batch(states[valid], values[valid].to(torch.float32))
        return int(valid.sum().item())
