"""Sim-side policy registry (crowd_sim/envs/policy/policy_factory.py:1-11)."""
from aemcarl_b200.crowd_sim.envs.policy.linear import Linear
from aemcarl_b200.crowd_sim.envs.policy.orca import ORCA


def none_policy():
    return None


policy_factory = dict()
policy_factory["linear"] = Linear
policy_factory["orca"] = ORCA
policy_factory["none"] = none_policy
