import sys, numpy as np, torch
sys.path.insert(0, '.')
from aemcarl_b200 import scenarios, weights as wts
from aemcarl_b200.engine import Engine, make_env_params
from aemcarl_b200.policy_math import build_action_space
eng = Engine(0); eng.set_action_space(build_action_space(1.0)); eng.set_env_params(make_env_params())
eng.load_weights(wts.synthetic_weights(0)); eng.set_precision(sys.argv[1])
B, n = 4096, 5
robot, humans = scenarios.generate_pool("train", 0, B, n, "circle_crossing")
hnext = np.concatenate([humans[:, :, 0:2], humans[:, :, 2:5]], axis=2)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for _ in range(2): eng.lookahead(d(robot), d(hnext), d(np.zeros((B, 81))), 1.0)
torch.cuda.synchronize()
