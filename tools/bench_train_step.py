"""Optimisation-step throughput of the value network (batch 100, n = 5, Adam): eager PyTorch against the CUDA-graph
replayed step.  usage: python tools/bench_train_step.py [--steps 300]"""
import argparse, copy, json, sys, time
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from aemcarl_b200.crowd_nav.utils.trainer import Trainer
from test_gpu_api import setup
ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=300); args = ap.parse_args()
env, robot, policy = setup()
dev = torch.device("cuda:0")
x = torch.randn((100, 5, 13), device=dev); v = torch.randn((100, 1), device=dev)
res = {}
for name, graph in (("eager", False), ("cuda_graph", True)):
    model = copy.deepcopy(policy.get_model()).to(dev)
    tr = Trainer(model, None, dev, 100, cuda_graph=graph)
    tr.set_learning_rate(1e-3, "Adam")
    for _ in range(5): tr._step(x, v)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(args.steps): loss = tr._step(x, v)
    float(loss); torch.cuda.synchronize()
    res[name] = {"steps_per_s": args.steps / (time.time() - t0), "ms_per_step": 1e3 * (time.time() - t0) / args.steps}
print(json.dumps(res))
