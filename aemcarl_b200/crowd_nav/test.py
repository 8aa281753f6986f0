"""Evaluation entry point mirroring crowd_nav/test.py:14-170: runs the seeded test cases and logs the reference's
summary line.  --vectorised evaluates all cases at once on the GPU (VecCrowdSim); the default plays them one by one through
the drop-in Explorer.  Unlike the reference (which hard-codes 50 episodes, test.py:169-170) --cases defaults to all 500."""
import argparse
import logging

import torch

from aemcarl_b200.crowd_nav.default_configs import env_config, policy_config
from aemcarl_b200.crowd_nav.policy.policy_factory import policy_factory
from aemcarl_b200.crowd_nav.utils.explorer import Explorer, run_vectorised_test
from aemcarl_b200.crowd_sim import make
from aemcarl_b200.crowd_sim.envs.utils.robot import Robot


def main(argv=None):
    ap = argparse.ArgumentParser("Parse configuration file")
    ap.add_argument("--model_weights", default=None, help="state_dict (.pth) of ValueNetworkTransformerAndGRU")
    ap.add_argument("--human_num", type=int, default=5)
    ap.add_argument("--square", action="store_true")
    ap.add_argument("--visible", action="store_true")
    ap.add_argument("--cases", type=int, default=500)
    ap.add_argument("--vectorised", action="store_true")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32x3", "fp16x3"])
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s, %(levelname)s: %(message)s")
    device = torch.device("cuda:0")
    ecfg = env_config(sim__human_num=args.human_num)
    policy = policy_factory["actenvcarl"]()
    policy.configure(policy_config())
    if args.model_weights:
        policy.get_model().load_state_dict(torch.load(args.model_weights, map_location="cpu"))
    policy.set_phase("test")
    policy.set_device(device)
    env = make("CrowdSim-v0")
    env.configure(ecfg)
    if args.square:
        env.test_sim = "square_crossing"
    robot = Robot(ecfg, "robot")
    robot.visible = args.visible
    robot.set_policy(policy)
    env.set_robot(robot)
    policy.set_env(env)
    if args.vectorised:
        from aemcarl_b200.vec_env import VecCrowdSim
        robot.time_step = env.time_step
        policy.time_step = env.time_step
        policy.build_action_space(robot.v_pref)
        eng = policy.engine()
        eng.set_env_params(_params(env, robot))
        eng.set_precision(args.precision)
        venv = VecCrowdSim(eng, min(args.cases, 4096), args.human_num, env.test_sim, "test", pool_size=args.cases,
                           gamma=policy.gamma)
        out = run_vectorised_test(venv, args.cases)
        logging.info("TEST  has success rate: %.2f, collision rate: %.2f, timeouts: %d, env-steps: %d",
                     out["success"] / args.cases, out["collision"] / args.cases, out["timeout"], out["env_steps"])
        return out
    explorer = Explorer(env, robot, device, gamma=policy.gamma)
    return explorer.run_k_episodes(args.cases, "test", print_failure=True)


def _params(env, robot):
    env.robot = robot
    env.humans = []
    return env.env_params()


if __name__ == "__main__":
    main()
