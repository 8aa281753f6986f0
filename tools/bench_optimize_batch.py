"""Trainer.optimize_batch on the on-device replay memory (100,000 pairs, batch 100): mini-batch sampling + the optimisation step:
the CUDA-graph replay of the torch step, the hand-written kernels (one C call per iteration), and the CUDA-graph replay of that call.
usage: python tools/bench_optimize_batch.py [--steps 2000]"""
import argparse, copy, json, sys, time
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from aemcarl_b200.crowd_nav.utils.trainer import Trainer
from aemcarl_b200.crowd_nav.utils.vec_explorer import DeviceReplayMemory
from test_gpu_api import setup
ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=2000); args = ap.parse_args()
env, robot, policy = setup()
dev = torch.device("cuda:0")
mem = DeviceReplayMemory(100000, 5, dev)
mem.push_batch(torch.randn((100000, 5, 13), device=dev), torch.randn((100000, 1), device=dev))
res = {}
for name, graph, fused in (("cuda_graph", True, False), ("fused_kernels", False, True), ("fused_kernels_graph", True, True)):
    model = copy.deepcopy(policy.get_model()).to(dev)
    tr = Trainer(model, mem, dev, 100, cuda_graph=graph, fused=fused)
    tr.set_learning_rate(1e-3, "Adam")
    tr.optimize_batch(20)
    torch.cuda.synchronize(); t0 = time.time()
    tr.optimize_batch(args.steps)
    torch.cuda.synchronize()
    res[name] = {"ms_per_step": 1e3 * (time.time() - t0) / args.steps}
torch.cuda.synchronize(); t0 = time.time()
for _ in range(args.steps): mem.sample(100)
torch.cuda.synchronize()
res["sample_only_ms"] = 1e3 * (time.time() - t0) / args.steps
print(json.dumps(res))
