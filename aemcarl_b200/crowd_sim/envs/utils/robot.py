from aemcarl_b200.crowd_sim.envs.utils.agent import Robot  # noqa: F401  (crowd_sim/envs/utils/robot.py)
