"""Optimisation-step throughput of the value network (batch 100, n = 5, Adam): eager PyTorch, the CUDA-graph replayed step
and the hand-written kernels of csrc/train.cu.  usage: python tools/bench_train_step.py [--steps 300]"""
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
for name, graph, fused in (("eager", False, False), ("cuda_graph", True, False), ("fused_kernels", False, True)):
    model = copy.deepcopy(policy.get_model()).to(dev)
    tr = Trainer(model, None, dev, 100, cuda_graph=graph, fused=fused)
    tr.set_learning_rate(1e-3, "Adam")
    for _ in range(5): tr._step(x, v)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(args.steps): loss = tr._step(x, v)
    float(loss); torch.cuda.synchronize()
    res[name] = {"steps_per_s": args.steps / (time.time() - t0), "ms_per_step": 1e3 * (time.time() - t0) / args.steps}
# device time of the fused step's four launches alone (CUDA events, no host gaps)
fs = tr._fused
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t = v.reshape(-1).contiguous()
e0.record()
for _ in range(args.steps):
    fs.engine.train_grad(fs.flat, x, t, fs.grad, fs.loss, fs.values, fs.steps)
e1.record(); torch.cuda.synchronize()
res["fused_kernels"]["train_grad_device_us"] = 1e3 * e0.elapsed_time(e1) / args.steps
res["fused_kernels"]["act_steps"] = int(fs.steps[0])
print(json.dumps(res))
