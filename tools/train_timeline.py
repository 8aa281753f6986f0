"""Phase timeline of train_fwdbwd_kernel (CTA 0): build with AEM_TRAIN_TIMELINE=1 python -m aemcarl_b200.build --force, then
python tools/train_timeline.py -- prints the cycles between consecutive barrier phases of the second call."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, ".")
from aemcarl_b200 import weights as wts, _lib
from aemcarl_b200.engine import Engine
eng = Engine(0); dev = eng.device
lib = _lib.load()
lib.aem_debug_train_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]; lib.aem_debug_train_timeline.restype = ctypes.c_int
p = torch.from_numpy(wts.synthetic_weights(3, ponder_bias=-2.0)).to(dev)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((100, 5, 13), device=dev, generator=g); t = torch.randn((100,), device=dev, generator=g)
grad = torch.empty_like(p)
buf = np.zeros(2048, dtype=np.int64)
for it in range(3):
    eng.train_grad(p, x, t, grad)
    n = lib.aem_debug_train_timeline(buf.ctypes.data_as(ctypes.c_void_p), 2048)
v = buf[:n]
t = np.abs(v)
d = np.diff(t)
jump = int(np.nonzero(d > 40000)[0][0]) if (d > 40000).any() else int(np.argmax(d))   # kernel boundary: probe stamps first
main_t, main_tag = t[jump + 1:], v[jump + 1:] < 0
print("main kernel: %d stamps, %d cycles" % (len(main_t), main_t[-1] - main_t[0]))
out = []
for i in range(1, len(main_t)):
    out.append(("w%d" if main_tag[i] else "p%d") % (main_t[i] - main_t[i - 1]))
print("p = cycles up to a barrier, w = cycles up to 'weights arrived' (mbarrier wait in take()):")
print(" ".join(out))
