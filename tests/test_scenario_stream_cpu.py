"""CPU: the per-thread code of csrc/scenario.cu (MT19937 stream, draw order, rejection loops, FP64 arithmetic) compiled as host
C++ and compared with the Python generator, which reproduces 208 cases recorded from the reference bit for bit
(tests/test_oracle_golden.py).  The GPU run of the same kernel is checked in tests/test_gpu_api.py; this test pins its logic
without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from aemcarl_b200 import scenarios
from aemcarl_b200._lib import ScenarioParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_kernel(tmp_path_factory):
    src = open(os.path.join(ROOT, "aemcarl_b200", "csrc", "scenario.cu")).read()
    body = src[src.index("struct Mt19937"):src.index("cudaError_t launch_scenarios")]
    body = (body.replace("__device__ __forceinline__", "static inline").replace("__device__", "")
            .replace("__global__ void __launch_bounds__(64) scenario_kernel(", "void scenario_host(int c, ")
            .replace("    const int c = blockIdx.x * blockDim.x + threadIdx.x;\n    if (c >= count) return;\n", "")
            .replace("__restrict__", "").replace("__longlong_as_double(0x7ff8000000000000LL)", "NAN"))
    d = tmp_path_factory.mktemp("scn")
    cpp = d / "scn.cpp"
    cpp.write_text("#include <cstdint>\n#include <cmath>\n"
                   "typedef struct { double circle_radius, square_width, discomfort_dist, robot_radius, robot_v_pref, human_radius, "
                   "human_v_pref; int32_t randomize_attributes, pad_; } aem_scenario_params;\n" + body +
                   'extern "C" void gen(int rule, uint32_t seed, int count, int n, aem_scenario_params *p, double *r, double *h, '
                   'int32_t *na)\n'
                   "{ for (int c = 0; c < count; ++c) scenario_host(c, rule, seed, count, n, *p, r, h, na); }\n")
    so = d / "scn.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(cpp)], check=True, timeout=120)
    return ctypes.CDLL(str(so))


@pytest.mark.timeout(120)
@pytest.mark.parametrize("rule,rid,n,randomize", [("square_crossing", 1, 10, False), ("square_crossing", 1, 5, True),
                                                  ("circle_crossing", 0, 5, False), ("circle_crossing", 0, 5, True),
                                                  ("circle_crossing", 0, 10, False)])
def test_kernel_logic_replays_the_reference_stream(host_kernel, rule, rid, n, randomize):
    count = 200
    for phase, first in (("train", 7), ("test", 0)):
        p = ScenarioParams(4.0, 10.0, 0.2, 0.3, 1.0, 0.3, 1.0, int(randomize), 0)
        r = np.empty((count, 9))
        h = np.empty((count, n, 8))
        host_kernel.gen(rid, ctypes.c_uint32(scenarios.case_seed(phase, first)), count, n, ctypes.byref(p),
                        r.ctypes.data_as(ctypes.c_void_p), h.ctypes.data_as(ctypes.c_void_p), None)
        rr, hh = scenarios.generate_pool(phase, first, count, n, rule, randomize_attributes=randomize)
        assert np.array_equal(r, rr)
        if rule == "square_crossing":
            assert np.array_equal(h, hh)                      # pure arithmetic: bit for bit
        else:
            assert np.abs(h - hh).max() <= 2 * np.spacing(4.0)   # libm cos / sin against numpy's
            assert np.array_equal(h[:, :, [2, 3, 4, 7]], hh[:, :, [2, 3, 4, 7]])


@pytest.mark.timeout(120)
@pytest.mark.parametrize("randomize", [False, True])
def test_kernel_logic_of_the_mixed_rule(host_kernel, randomize):
    """rule 2 (crowd_sim.py:337-386): static / dynamic decision, human count, static obstacles, padding rows and counts"""
    count, n = 400, scenarios.MIXED_MAX_HUMANS
    for phase, first in (("train", 3), ("test", 100)):
        p = ScenarioParams(4.0, 10.0, 0.2, 0.3, 1.0, 0.3, 1.0, int(randomize), 0)
        r = np.empty((count, 9))
        h = np.empty((count, n, 8))
        na = np.zeros(count, dtype=np.int32)
        host_kernel.gen(2, ctypes.c_uint32(scenarios.case_seed(phase, first)), count, n, ctypes.byref(p),
                        r.ctypes.data_as(ctypes.c_void_p), h.ctypes.data_as(ctypes.c_void_p), na.ctypes.data_as(ctypes.c_void_p))
        rr, hh, cc = scenarios.generate_mixed_pool(phase, first, count, randomize_attributes=randomize)
        assert np.array_equal(na, cc) and np.array_equal(r, rr)
        assert set(cc.tolist()) == {1, 2, 3, 4, 5}
        assert np.array_equal(h[:, 2:], hh[:, 2:])                   # square-crossing humans, static ones, padding rows
        assert np.abs(h - hh).max() <= 2 * np.spacing(4.0)           # circle-crossing humans: libm cos / sin against numpy's
        static = np.array([(hh[i, :cc[i], 0:2] == hh[i, :cc[i], 5:7]).all() for i in range(count)])
        assert static.sum() > 40 and np.array_equal(h[static], hh[static])
