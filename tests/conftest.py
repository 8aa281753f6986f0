import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def actions():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "action_rotate.npz"))["actions"]


# The default GPU suite runs every lookahead test under the FP32 anchor kernel and the shipped FP16x3 kernel;
# AEM_TEST_ALL_PRECISIONS=1 adds the 3xTF32 kernel (FP32-range containers, kept as an alternative).
PRECISIONS = ["fp32", "fp16x3"] + (["tf32x3"] if os.environ.get("AEM_TEST_ALL_PRECISIONS") else [])


@pytest.fixture(scope="session", params=PRECISIONS)
def engine(actions, request):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from aemcarl_b200.engine import Engine, make_env_params
    eng = Engine(0)
    eng.set_action_space(actions)
    eng.set_env_params(make_env_params())
    eng.set_precision(request.param)
    eng.precision_name = request.param
    return eng
