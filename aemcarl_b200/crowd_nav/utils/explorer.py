"""Episode driver (crowd_nav/utils/explorer.py:22-151): runs k episodes, gathers the outcome statistics, fills the
replay memory with (rotated state, target value) pairs.  Same log-line format (utils/plot.py:38-55 parses it)."""
import copy
import logging
import time as t

import torch

from aemcarl_b200.crowd_sim.envs.utils.info import Collision, Danger, ReachGoal, Timeout


def average(input_list):
    return sum(input_list) / len(input_list) if input_list else 0


def summary_line(phase, episode, success_rate, collision_rate, avg_nav_time, total_reward):
    """The per-run summary of explorer.py:98-100, character for character: crowd_nav/utils/plot.py:38-55 parses the
    TRAIN / VAL lines of output.log with regular expressions."""
    extra_info = "" if episode is None else "in episode {} ".format(episode)
    return "{:<5} {}has success rate: {:.2f}, collision rate: {:.2f}, nav time: {:.2f}, total reward: {:.4f}".format(
        phase.upper(), extra_info, success_rate, collision_rate, avg_nav_time, total_reward)


class Explorer(object):
    def __init__(self, env, robot, device, memory=None, gamma=None, target_policy=None):
        self.env = env
        self.robot = robot
        self.device = device
        self.memory = memory
        self.gamma = gamma
        self.target_policy = target_policy
        self.target_model = None
        self.last_stats = None

    def update_target_model(self, target_model):
        self.target_model = copy.deepcopy(target_model)

    def run_k_episodes(self, k, phase, update_memory=False, imitation_learning=False, episode=None,
                       print_failure=False, verbose=False):
        self.robot.policy.set_phase(phase)
        success_times, collision_times, timeout_times = [], [], []
        success = collision = timeout = too_close = 0
        min_dist, cumulative_rewards, collision_cases, timeout_cases = [], [], [], []
        outcomes = []
        for i in range(k):
            time_begin = t.time()
            ob = self.env.reset(phase)
            done = False
            states, actions, rewards = [], [], []
            while not done:
                action = self.robot.act(ob)
                ob, reward, done, info = self.env.step(action)
                states.append(self.robot.policy.last_state)
                actions.append(action)
                rewards.append(reward)
                if isinstance(info, Danger):
                    too_close += 1
                    min_dist.append(info.min_dist)
            if isinstance(info, ReachGoal):
                success += 1
                success_times.append(self.env.global_time)
            elif isinstance(info, Collision):
                collision += 1
                collision_cases.append(i)
                collision_times.append(self.env.global_time)
            elif isinstance(info, Timeout):
                timeout += 1
                timeout_cases.append(i)
                timeout_times.append(self.env.time_limit)
            else:
                raise ValueError("Invalid end signal from environment")
            outcomes.append(str(info))
            if update_memory and (isinstance(info, ReachGoal) or isinstance(info, Collision)):
                self.update_memory(states, actions, rewards, imitation_learning)
            cumulative_rewards.append(sum(pow(self.gamma, s * self.robot.time_step * self.robot.v_pref) * reward
                                          for s, reward in enumerate(rewards)))
            if verbose:
                print("i/k: %d/%d, robot pos: %s, %s, time consuming is:%s secs. %s " %
                      (i, k, self.robot.px, self.robot.py, t.time() - time_begin, str(info)))
        if getattr(self.robot.policy, "use_rate_statistic", False) and verbose:
            print("ACT Usage: ", self.robot.policy.get_act_statistic_info())
        success_rate = success / k
        collision_rate = collision / k
        assert success + collision + timeout == k
        avg_nav_time = sum(success_times) / len(success_times) if success_times else self.env.time_limit
        logging.info(summary_line(phase, episode, success_rate, collision_rate, avg_nav_time,
                                  average(cumulative_rewards)))
        if phase in ["val", "test"]:
            total_time = sum(success_times + collision_times + timeout_times) * self.robot.time_step
            logging.info("Frequency of being in danger: %.2f and average min separate distance in danger: %.2f",
                         too_close / total_time, average(min_dist))
        if print_failure:
            logging.info("Collision cases: " + " ".join([str(x) for x in collision_cases]))
            logging.info("Timeout cases: " + " ".join([str(x) for x in timeout_cases]))
        self.last_stats = dict(success=success, collision=collision, timeout=timeout, too_close=too_close,
                               avg_nav_time=avg_nav_time, cumulative_rewards=cumulative_rewards, outcomes=outcomes)
        return self.last_stats

    def update_memory(self, states, actions, rewards, imitation_learning=False):
        if self.memory is None or self.gamma is None:
            raise ValueError("Memory or gamma value is not set!")
        for i, state in enumerate(states):
            reward = rewards[i]
            if imitation_learning:
                state = self.target_policy.transform(state)
                value = sum(pow(self.gamma, max(s - i, 0) * self.robot.time_step * self.robot.v_pref) * r
                            * (1 if s >= i else 0) for s, r in enumerate(rewards))
            else:
                if i == len(states) - 1:
                    value = reward
                else:
                    next_state = states[i + 1]
                    gamma_bar = pow(self.gamma, self.robot.time_step * self.robot.v_pref)
                    with torch.no_grad():
                        action_value, _ = self.target_model(next_state.unsqueeze(0))
                    value = reward + gamma_bar * action_value.data.item()
            value = torch.Tensor([value]).to(self.device)
            self.memory.push((state, value))


def run_vectorised_test(vec_env, num_cases):
    """Evaluate scenarios 0 .. num_cases-1 of `vec_env`'s pool, B at a time, one episode each (the vectorised form of
    run_k_episodes(k, 'test')).  Returns the outcome counts, the per-case info codes and the env-steps executed."""
    import numpy as np
    B = vec_env.B
    assert num_cases <= vec_env.pool_size
    limit = int(round(vec_env.engine.params.time_limit / vec_env.engine.params.time_step)) + 2
    codes = np.full(num_cases, -1, dtype=np.int64)
    steps = 0
    for first in range(0, num_cases, B):
        vec_env.load_cases(first)
        live = np.arange(B) + first < num_cases
        pending = live.copy()
        for _ in range(limit):
            reward, done, info, dmin = vec_env.decide_and_step(auto_reset=False)
            steps += int(pending.sum())
            d = done.cpu().numpy() != 0
            newly = pending & d
            codes[first + np.nonzero(newly)[0]] = info.cpu().numpy()[newly]
            pending &= ~d
            if not pending.any():
                break
    return dict(success=int((codes == 2).sum()), collision=int((codes == 3).sum()), timeout=int((codes == 4).sum()),
                codes=codes, env_steps=steps)
