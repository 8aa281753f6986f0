"""GPU: the north-star check with TRAINED weights -- a short training run (imitation + vectorised RL through the fused
optimisation step, ~13 s: the run of tools/train_demo.py), then the 500 seeded test cases: the batched GPU pipeline against the CPU oracle pipeline on identical
scenarios, for the FP32 and the FP16x3 lookahead kernels.  (tests/test_gpu_episodes.py does the same with random-init weights,
whose episodes mostly time out; a trained policy reaches the goal, halts the adaptive-computation-time loop at different step
counts and drives through crowds, so the trajectories exercise the reward / collision / goal branches.)"""
import numpy as np
import pytest
import torch

from aemcarl_b200 import scenarios
from aemcarl_b200.crowd_nav.utils.explorer import run_vectorised_test
from aemcarl_b200.engine import Engine, make_env_params
from aemcarl_b200.vec_env import VecCrowdSim
from test_gpu_episodes import run_case_set

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def trained_flat(tmp_path_factory):
    from aemcarl_b200.crowd_nav import train
    out = train.main(["--no_final_test", "--output_dir", str(tmp_path_factory.mktemp("trained")), "--il_episodes", "2048", "--il_epochs", "20",
                      "--train_episodes", "16384", "--vectorised", "2048", "--updates_per_wave", "2000", "--fused_step"])
    return out["model"].flat_weights().copy()


def test_trained_policy_outcomes_match_oracle(oracle, actions, trained_flat):
    """500 / 500 identical outcomes (or every differing case traced to a near-tie decision by the lock-step probe)."""
    for precision in ("fp32", "fp16x3"):
        out, codes, probe = run_case_set(oracle, actions, trained_flat, precision, 500, 5, "circle_crossing")
        assert (codes == 2).mean() > 0.5, "the short training run should already reach the goal in most cases"
