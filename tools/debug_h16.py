import sys, numpy as np, torch
sys.path.insert(0, '.')
from aemcarl_b200 import scenarios, weights as wts
from aemcarl_b200.engine import Engine, make_env_params
from aemcarl_b200.policy_math import build_action_space
eng = Engine(0); eng.set_action_space(build_action_space(1.0)); eng.set_env_params(make_env_params())
eng.load_weights(wts.synthetic_weights(0))
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
for B in (64, 100, 300, 512, 4096):
    n = 5
    robot, humans = scenarios.generate_pool("train", 0, B, n, "circle_crossing")
    rng = np.random.RandomState(0)
    humans[:, :, 2:4] = rng.uniform(-0.7, 0.7, (B, n, 2)); robot[:, 0:2] += rng.uniform(-1, 1, (B, 2))
    hnext = np.concatenate([humans[:, :, 0:2] + 0.2 * humans[:, :, 2:4], humans[:, :, 2:5]], axis=2)
    rewards = np.zeros((B, 81))
    args = (dev(robot), dev(hnext), dev(rewards), 1.0)
    eng.set_precision("fp32"); vr = eng.lookahead(*args)[1].cpu().numpy().ravel()
    eng.set_precision(sys.argv[1] if len(sys.argv) > 1 else "fp16x3"); v = eng.lookahead(*args)[1].cpu().numpy().ravel()
    rel = np.abs(v - vr) / np.maximum(np.abs(vr), 1e-2)
    bad = np.nonzero(rel > 1e-4)[0]
    S = 21; ntiles = (B * 81 + S - 1) // S
    print("B", B, "tiles", ntiles, "max rel %.2e" % rel.max(), "bad", len(bad))
    if len(bad):
        tiles = bad // S
        print("  bad tiles (first 20):", np.unique(tiles)[:20], " tile%296:", np.unique(tiles % 296)[:20], " n bad tiles", len(np.unique(tiles)))
        print("  pos in tile:", np.bincount(bad % S, minlength=S))
        print("  tile//296 hist:", np.bincount(np.unique(tiles) // 296))
