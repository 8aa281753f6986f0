"""Default configurations as RawConfigParser objects.

The reference drives everything from three INI files (crowd_nav/configs/{env,policy,train}.config) read with
configparser.RawConfigParser; this package reads the SAME sections / keys, so a user's existing files work
unchanged (`config.read(path)`).  These helpers only provide the reference's default values for the hot path
without shipping copies of the files.
"""
import configparser

_ENV = {
    "env": dict(time_limit=20, time_step=0.2, val_size=100, test_size=500, randomize_attributes="false"),
    "reward": dict(success_reward=1, collision_penalty=-0.25, discomfort_dist=0.2, discomfort_penalty_factor=0.5,
                   stationary_penalty=-0.3, stationary_penalty_dist=0.2, agent_timestep=0.4, human_timestep=0.6,
                   reward_increment=4.0, position_variance=4.0, direction_variance=3.0),
    "sim": dict(train_val_sim="circle_crossing", test_sim="circle_crossing", square_width=10, circle_radius=4.0,
                human_num=5),
    "humans": dict(visible="true", policy="orca", radius=0.3, v_pref=1, sensor="coordinates"),
    "robot": dict(visible="false", policy="none", radius=0.3, v_pref=1.0, sensor="coordinates"),
}

_POLICY = {
    "rl": dict(gamma=0.9),
    "om": dict(cell_num=4, cell_size=1, om_channel_size=3),
    "action_space": dict(kinematics="holonomic", speed_samples=5, rotation_samples=16, sampling="exponential",
                         query_env="true"),
    "actcarl": dict(multiagent_training="true"),
    "actenvcarl": dict(in_mlp_dims="100, 50", sort_mlp_dims="100, 50", sort_attention_dims="50, 50, 1",
                       action_dims="100, 50, 1", global_state_dim=50, with_dynamic_net="true",
                       multiagent_training="true", with_global_state="false", with_om="false",
                       test_policy_flag="5", multi_process="self_attention", act_steps=1, act_fixed=False),
}

_TRAIN = {
    "trainer": dict(batch_size=100),
    "imitation_learning": dict(il_episodes=3000, il_policy="orca", il_epochs=50, il_learning_rate=0.01,
                               safety_space=0.15),
    "train": dict(rl_learning_rate=0.001, train_batches=100, train_episodes=10000, sample_episodes=1,
                  target_update_interval=50, evaluation_interval=1000, capacity=100000, epsilon_start=0.5,
                  epsilon_end=0.1, epsilon_decay=4000, checkpoint_interval=200),
}


def _make(d, overrides):
    cfg = configparser.RawConfigParser()
    for sec, kv in d.items():
        cfg.add_section(sec)
        for k, v in kv.items():
            cfg.set(sec, k, v if isinstance(v, (bool, int)) and sec == "actenvcarl" and k in ("act_steps", "act_fixed")
                    else str(v))
    for (sec, k), v in (overrides or {}).items():
        cfg.set(sec, k, v)
    return cfg


def env_config(readme_reward_flags=True, **overrides):
    """env.config defaults; readme_reward_flags applies the CLI values of README.md:56 / test.py:36-40
    (agent_timestep 0.4, human_timestep 0.5, reward_increment 2.0, position/direction_variance 2.0)."""
    cfg = _make(_ENV, None)
    if readme_reward_flags:
        for k, v in dict(agent_timestep=0.4, human_timestep=0.5, reward_increment=2.0, position_variance=2.0,
                         direction_variance=2.0).items():
            cfg.set("reward", k, str(v))
    for k, v in overrides.items():
        sec, key = k.split("__")
        cfg.set(sec, key, v if not isinstance(v, (int, float)) or isinstance(v, bool) else str(v))
    return cfg


def policy_config(act_fixed=False, act_steps=1):
    cfg = _make(_POLICY, None)
    cfg.set("actenvcarl", "act_fixed", act_fixed)      # raw Python objects, as train.py:118-121 injects them
    cfg.set("actenvcarl", "act_steps", act_steps)
    return cfg


def train_config():
    return _make(_TRAIN, None)
