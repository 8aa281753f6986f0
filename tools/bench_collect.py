"""Experience-collection throughput of the RL phase: the serial drop-in Explorer (one episode at a time,
explorer.py:44-80) against the vectorised explorer with the on-device replay memory.  Prints one JSON line.
usage: python tools/bench_collect.py [--envs 1024] [--episodes 4096] [--serial_episodes 8]"""
import argparse, json, sys, time
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from aemcarl_b200 import weights as wts
from aemcarl_b200.crowd_nav.utils.explorer import Explorer
from aemcarl_b200.crowd_nav.utils.memory import ReplayMemory
from aemcarl_b200.crowd_nav.utils.vec_explorer import DeviceReplayMemory, VecExplorer
from test_gpu_api import setup

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1024)
ap.add_argument("--episodes", type=int, default=4096)
ap.add_argument("--serial_episodes", type=int, default=8)
ap.add_argument("--epsilon", type=float, default=0.1)
ap.add_argument("--precision", default="fp16x3")
args = ap.parse_args()
env, robot, policy = setup()
dev = torch.device("cuda:0")
policy.set_phase("train")
policy.set_epsilon(args.epsilon)
policy.engine().set_precision(args.precision)
mem = ReplayMemory(100000)
serial = Explorer(env, robot, dev, mem, policy.gamma, policy)
serial.update_target_model(policy.get_model())
env.case_counter["train"] = 0
serial.run_k_episodes(1, "train", update_memory=True)          # warm-up
torch.cuda.synchronize(); t0 = time.time()
steps0 = 0
orig_step = env.step
def counting_step(*a, **k):
    global steps0
    steps0 += 1
    return orig_step(*a, **k)
env.step = counting_step
serial.run_k_episodes(args.serial_episodes, "train", update_memory=True)
torch.cuda.synchronize(); t_serial = time.time() - t0

dmem = DeviceReplayMemory(1000000, 5, dev)
vec = VecExplorer(policy, args.envs, env.env_params(), dmem, policy.gamma, human_num=5)
vec.update_target_model(policy.get_model())
vec.run_k_episodes(args.envs, "train", update_memory=True, epsilon=args.epsilon, first_case=0)     # warm-up
dmem.clear()
torch.cuda.synchronize(); t0 = time.time()
st = vec.run_k_episodes(args.episodes, "train", update_memory=True, epsilon=args.epsilon, first_case=10000)
torch.cuda.synchronize(); t_vec = time.time() - t0
print(json.dumps({"serial": {"episodes": args.serial_episodes, "env_steps": steps0, "seconds": t_serial,
                             "env_steps_per_s": steps0 / t_serial, "episodes_per_s": args.serial_episodes / t_serial},
                  "vectorised": {"envs": args.envs, "episodes": args.episodes, "env_steps": st["env_steps"], "seconds": t_vec,
                                 "env_steps_per_s": st["env_steps"] / t_vec, "episodes_per_s": args.episodes / t_vec,
                                 "pairs_pushed": st["pushed"], "outcomes": [st["success"], st["collision"], st["timeout"]]},
                  "precision": args.precision, "epsilon": args.epsilon}))
