from aemcarl_b200.crowd_sim.envs.crowd_sim import CrowdSim  # noqa: F401
