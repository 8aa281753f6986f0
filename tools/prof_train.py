"""One training-gradient call of csrc/train.cu at the trainer's shape (batch 100, 5 humans) for ncu:
    ncu --set full --clock-control none --import-source on -k regex:train_fwdbwd -s 2 -c 1 -o gpurun_out/trainprof -f \
        python tools/prof_train.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from aemcarl_b200 import weights as wts
from aemcarl_b200.engine import Engine
eng = Engine(0)
dev = eng.device
pb = float(sys.argv[1]) if len(sys.argv) > 1 else -2.0          # -2: three GRU steps; 4: one
p = torch.from_numpy(wts.synthetic_weights(3, ponder_bias=pb)).to(dev)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((100, 5, 13), device=dev, generator=g)
t = torch.randn((100,), device=dev, generator=g)
grad = torch.empty_like(p)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(4):
    out = eng.train_grad(p, x, t, grad)
torch.cuda.synchronize()
e0.record()
for _ in range(50):
    out = eng.train_grad(p, x, t, grad)
e1.record(); torch.cuda.synchronize()
print("train_grad: %.1f us per call, K = %d, loss %.6f" % (1e3 * e0.elapsed_time(e1) / 50, int(out[3][0]), float(out[1][0])))
