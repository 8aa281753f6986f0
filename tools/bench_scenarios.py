"""Scenario generation: host generator (bit-exact mirror of the reference) against aem_generate_scenarios on the device."""
import json, sys, time
import torch
sys.path.insert(0, ".")
from aemcarl_b200 import scenarios
from aemcarl_b200.engine import Engine
eng = Engine(0)
res = {}
for rule, n in (("circle_crossing", 5), ("square_crossing", 10)):
    t0 = time.time(); scenarios.generate_pool("train", 0, 1024, n, rule); host_ms_per_1k = 1e3 * (time.time() - t0)
    for count in (4096, 65536):
        eng.generate_scenarios("train", 0, count, n, rule); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(5): eng.generate_scenarios("train", 1000 * i, count, n, rule)
        e1.record(); torch.cuda.synchronize()
        res["%s_n%d_%d" % (rule, n, count)] = {"device_ms": e0.elapsed_time(e1) / 5, "host_ms_extrapolated": host_ms_per_1k * count / 1024}
print(json.dumps(res))
