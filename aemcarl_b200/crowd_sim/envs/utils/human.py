from aemcarl_b200.crowd_sim.envs.utils.agent import Human  # noqa: F401  (crowd_sim/envs/utils/human.py)
