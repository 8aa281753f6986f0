"""Policy registry (crowd_nav/policy/policy_factory.py:12-19).  Only the north-star policy is trainable here; the
reference's ablation policies (cadrl, lstm_rl, sarl, comcarl, gipcarl, actcarl, actfcarl) are out of scope."""
from aemcarl_b200.crowd_sim.envs.policy.policy_factory import policy_factory
from aemcarl_b200.crowd_nav.policy.actenvcarl import ACTENVCARL

policy_factory["actenvcarl"] = ACTENVCARL
