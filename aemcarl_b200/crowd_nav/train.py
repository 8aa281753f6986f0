"""Training entry point mirroring crowd_nav/train.py:20-246 for the actenvcarl policy on the CUDA hot path.

    python -m aemcarl_b200.crowd_nav.train --output_dir data/output [--il_episodes 200] [--train_episodes N] ...
    torchrun --nproc-per-node N -m aemcarl_b200.crowd_nav.train ...     (data parallel: env shards + grad all-reduce)

Flow (same order as the reference): imitation learning with the ORCA teacher (safety_space from train.config) ->
Trainer.optimize_epoch -> il_model.pth -> target model -> RL loop with the linear epsilon schedule, one sampled episode
per step, optimize_batch, periodic target update / validation / checkpoints (rl_model_{ep}.pth, state_dict keys of the
reference).  Data parallelism: every rank plays its own cases (rank-strided case counters) and the flat gradient is
all-reduced between backward() and step() (crowd_nav/utils/trainer.py).
"""
import argparse
import copy
import logging
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

from aemcarl_b200.crowd_nav.default_configs import env_config, policy_config, train_config
from aemcarl_b200.crowd_nav.policy.policy_factory import policy_factory
from aemcarl_b200.crowd_nav.utils.explorer import Explorer
from aemcarl_b200.crowd_nav.utils.memory import ReplayMemory
from aemcarl_b200.crowd_nav.utils.trainer import Trainer, broadcast_parameters, min_over_ranks
from aemcarl_b200.crowd_sim import make
from aemcarl_b200.crowd_sim.envs.utils.robot import Robot


def build(args, device):
    ecfg = env_config(sim__human_num=args.human_num)
    for k in ("agent_timestep", "human_timestep", "reward_increment", "position_variance", "direction_variance"):
        ecfg.set("reward", k, str(getattr(args, k)))
    pcfg = policy_config(act_fixed=args.act_fixed, act_steps=args.act_steps)
    tcfg = train_config()
    policy = policy_factory["actenvcarl"]()
    policy.configure(pcfg)
    policy.set_device(device)
    env = make("CrowdSim-v0")
    env.configure(ecfg)
    robot = Robot(ecfg, "robot")
    env.set_robot(robot)
    return ecfg, pcfg, tcfg, policy, env, robot


def main(argv=None):
    ap = argparse.ArgumentParser("Parse configuration file")
    ap.add_argument("--output_dir", default="data/output")
    ap.add_argument("--human_num", type=int, default=5)
    ap.add_argument("--optimizer", default="Adam")
    ap.add_argument("--il_episodes", type=int, default=200)           # the reference hard-codes 200 (train.py:196)
    ap.add_argument("--il_epochs", type=int, default=None)
    ap.add_argument("--train_episodes", type=int, default=None)
    ap.add_argument("--agent_timestep", type=float, default=0.4)
    ap.add_argument("--human_timestep", type=float, default=0.5)
    ap.add_argument("--reward_increment", type=float, default=2.0)
    ap.add_argument("--position_variance", type=float, default=2.0)
    ap.add_argument("--direction_variance", type=float, default=2.0)
    ap.add_argument("--act_steps", type=int, default=1)
    ap.add_argument("--act_fixed", default=False, action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--resume", default=False, action="store_true")   # train.py:171-177
    ap.add_argument("--visible", default=False, action="store_true")  # train.py:46
    ap.add_argument("--dropout", action="store_true",
                    help="optimise in train() mode like the reference (it never calls eval(): the encoder layers' dropout, "
                         "p = 0.1 at 4 sites per layer, is live in every optimisation step; qnetwork_factory.py:628)")
    ap.add_argument("--target_dropout", action="store_true",
                    help="vectorised RL phase: evaluate the target network in train() mode as well (the reference's V_target "
                         "of explorer.py:132-139 is a stochastic forward because it never calls eval())")
    ap.add_argument("--no_final_test", action="store_true",
                    help="skip the reference's episode-0 validation and closing run_k_episodes(test) (train.py:229-230,246) "
                         "-- benchmarks / unit tests only")
    ap.add_argument("--vectorised", type=int, default=0, metavar="B",
                    help="RL phase on B environments in lock-step (VecExplorer + on-device replay memory); 0 = the "
                         "reference's one-episode-at-a-time loop")
    ap.add_argument("--cuda_graph", action="store_true",
                    help="replay the optimisation step (forward, MSE, backward, Adam) from CUDA graphs")
    ap.add_argument("--fused_step", action="store_true",
                    help="the optimisation step as hand-written kernels (csrc/train.cu: aem_train_grad + aem_adam_step)")
    ap.add_argument("--updates_per_wave", type=int, default=None,
                    help="optimizer steps after each wave of B episodes (default: train_batches x B, the reference's ratio)")
    ap.add_argument("--precision", default="fp16x3", choices=["fp32", "tf32x3", "fp16x3", "fp16x1"],
                    help="arithmetic of the lookahead kernel in the vectorised RL phase")
    args = ap.parse_args(argv)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s, %(levelname)s: %(message)s",
                        datefmt="%Y-%m-%d %H:%M:%S")
    os.makedirs(args.output_dir, exist_ok=True)
    torch.manual_seed(args.seed)
    ecfg, pcfg, tcfg, policy, env, robot = build(args, device)
    model = policy.get_model()
    broadcast_parameters(model, 0)

    batch_size = tcfg.getint("trainer", "batch_size")
    vec = None
    if args.vectorised > 0:      # experience lives in HBM from the start (utils/vec_explorer.py)
        from aemcarl_b200.crowd_nav.utils.vec_explorer import DeviceReplayMemory, VecExplorer
        memory = DeviceReplayMemory(tcfg.getint("train", "capacity"), args.human_num, device)
    else:
        memory = ReplayMemory(tcfg.getint("train", "capacity"))
    if args.dropout:
        model.train()             # the optimisation steps see dropout; the lookahead kernels act without it (documented)
    trainer = Trainer(model, memory, device, batch_size, cuda_graph=args.cuda_graph, fused=args.fused_step)
    explorer = Explorer(env, robot, device, memory, policy.gamma, target_policy=policy)

    # ---- imitation learning (train.py:171-204): --resume / an existing il_model.pth skip it, like the reference
    robot.visible = bool(args.visible)
    il_weight_file = os.path.join(args.output_dir, "il_model.pth")
    rl_weight_file = os.path.join(args.output_dir, "rl_model.pth")
    policy.set_env(env)
    policy.time_step = env.time_step
    loss = float("nan")
    if args.vectorised > 0:
        params = env.env_params()
        params.orca_safety_space = 0.0                    # the humans' own ORCA; the teacher's margin goes to VecExplorer
        policy.build_action_space(robot.v_pref)
    il_policy = policy_factory[tcfg.get("imitation_learning", "il_policy")]()
    il_policy.multiagent_training = policy.multiagent_training
    # the teacher keeps no extra margin when the humans see the robot (train.py:186-189)
    il_policy.safety_space = 0.0 if robot.visible else tcfg.getfloat("imitation_learning", "safety_space")
    if args.vectorised > 0:
        vec = VecExplorer(policy, args.vectorised, params, memory, policy.gamma, human_num=args.human_num,
                          rule=env.train_val_sim, seed=args.seed * 1000 + rank,
                          il_safety_space=il_policy.safety_space, target_train_mode=args.target_dropout)
    if args.resume:
        if not os.path.exists(rl_weight_file):
            logging.error("RL weights does not exist")
        model.load_state_dict(torch.load(rl_weight_file, map_location=device))
        rl_weight_file = os.path.join(args.output_dir, "resumed_rl_model.pth")
        logging.info("Load reinforcement learning trained weights. Resume training")
    elif os.path.exists(il_weight_file):
        model.load_state_dict(torch.load(il_weight_file, map_location=device))
        logging.info("Load imitation learning trained weights.")
    else:
        robot.set_policy(il_policy)
        trainer.set_learning_rate(tcfg.getfloat("imitation_learning", "il_learning_rate"), args.optimizer)
        if args.vectorised > 0:
            per_rank = (args.il_episodes + world - 1) // world
            vec.run_k_episodes(per_rank, "train", update_memory=True, imitation_learning=True, first_case=rank * per_rank)
            env.case_counter["train"] = per_rank * world
        else:
            env.case_counter["train"] = rank                  # rank-strided cases
            stride_cases(env, world)
            explorer.run_k_episodes(args.il_episodes, "train", update_memory=True, imitation_learning=True)
        il_epochs = args.il_epochs or tcfg.getint("imitation_learning", "il_epochs")
        # every rank must take the same all-reduced steps: the decision is made on the SMALLEST memory of all ranks
        if min_over_ranks(len(memory), device) >= batch_size:
            loss = trainer.optimize_epoch(il_epochs)
        logging.info("Finish imitation learning. Experience set size: %d/%d, loss %.3e", len(memory), memory.capacity, loss)
        if rank == 0:
            torch.save(model.state_dict(), il_weight_file)
    explorer.update_target_model(model)

    # ---- reinforcement learning (train.py:206-246)
    robot.set_policy(policy)
    policy.set_env(env)
    trainer.set_learning_rate(tcfg.getfloat("train", "rl_learning_rate"), args.optimizer)
    train_episodes = args.train_episodes if args.train_episodes is not None else tcfg.getint("train", "train_episodes")
    eps_start, eps_end = tcfg.getfloat("train", "epsilon_start"), tcfg.getfloat("train", "epsilon_end")
    eps_decay = tcfg.getfloat("train", "epsilon_decay")
    if args.resume:                                        # fill the memory with some RL experience (train.py:212-216)
        policy.set_epsilon(eps_end)
        if vec is not None:
            vec.update_target_model(model)
            vec.run_k_episodes(100, "train", update_memory=True, episode=0, epsilon=eps_end,
                               first_case=env.case_counter["train"] + rank * 100)
        else:
            explorer.run_k_episodes(100, "train", update_memory=True, episode=0)
        logging.info("Experience set size: %d/%d", len(memory), memory.capacity)
    sched = (eps_end, eps_end, 1.0) if args.resume else (eps_start, eps_end, eps_decay)
    episode = 0
    if args.vectorised > 0:
        episode = vectorised_rl(args, tcfg, policy, env, model, vec, trainer, explorer, train_episodes, sched, rank, world)
    while episode < train_episodes:
        eps_s, eps_e, eps_d = sched
        epsilon = eps_s + (eps_e - eps_s) / eps_d * episode if episode < eps_d else eps_e
        policy.set_epsilon(epsilon)
        if episode % tcfg.getint("train", "evaluation_interval") == 0 and not (episode == 0 and args.no_final_test):
            # episode 0 included (train.py:229-230)
            explorer.run_k_episodes(env.case_size["val"], "val", episode=episode)
        explorer.run_k_episodes(tcfg.getint("train", "sample_episodes"), "train", update_memory=True, episode=episode)
        if min_over_ranks(len(memory), device) >= 1:
            trainer.optimize_batch(tcfg.getint("train", "train_batches"))
        episode += 1
        if episode % tcfg.getint("train", "target_update_interval") == 0:
            explorer.update_target_model(model)
        if episode != 0 and episode % tcfg.getint("train", "checkpoint_interval") == 0 and rank == 0:
            torch.save(model.state_dict(), os.path.join(args.output_dir, "rl_model_%d.pth" % episode))
    if rank == 0:
        torch.save(model.state_dict(), rl_weight_file)
    # final test (train.py:246)
    if not args.no_final_test:
        if vec is not None:
            vec.run_k_episodes(env.case_size["test"], "test", episode=episode, first_case=0)
        else:
            explorer.run_k_episodes(env.case_size["test"], "test", episode=episode)
    return dict(loss=loss, memory=len(memory), model=model, explorer=explorer, policy=policy, env=env, robot=robot,
                trainer=trainer)


def vectorised_rl(args, tcfg, policy, env, model, vec, trainer, explorer, train_episodes, eps_sched, rank, world):
    """RL phase of train.py:206-246 with waves of B episodes per rank: same epsilon schedule (evaluated at the episode
    counter of the wave), same sample : update ratio, same target-update / validation / checkpoint intervals (triggered
    when a wave crosses a multiple).  Experience stays in the DeviceReplayMemory the imitation phase filled."""
    B = args.vectorised
    dmem = vec.memory
    policy.engine().set_precision(args.precision)
    vec.update_target_model(model)
    eps_start, eps_end, eps_decay = eps_sched
    base_case = env.case_counter["train"]
    updates = args.updates_per_wave or tcfg.getint("train", "train_batches") * B // tcfg.getint("train", "sample_episodes")
    every = lambda key, before, after: after // tcfg.getint("train", key) > before // tcfg.getint("train", key)
    episode, wave = 0, 0
    stats = None
    if train_episodes > 0 and not args.no_final_test:          # the reference validates at episode 0 too (train.py:229-230)
        vec.run_k_episodes(env.case_size["val"], "val", episode=0, first_case=0)
    while episode < train_episodes:
        epsilon = eps_start + (eps_end - eps_start) / eps_decay * episode if episode < eps_decay else eps_end
        policy.set_epsilon(epsilon)
        stats = vec.run_k_episodes(B, "train", update_memory=True, episode=episode, epsilon=epsilon,
                                   first_case=base_case + (wave * world + rank) * B)
        # ranks store different numbers of success / collision pairs: the all-reduced update runs on every rank or on none
        if min_over_ranks(len(dmem), trainer.device) >= trainer.batch_size:
            trainer.optimize_batch(updates)
        before, episode, wave = episode, episode + B * world, wave + 1
        if every("target_update_interval", before, episode):
            vec.update_target_model(model)
        if every("evaluation_interval", before, episode) and episode < train_episodes:
            vec.run_k_episodes(env.case_size["val"], "val", episode=episode, first_case=0)
        if every("checkpoint_interval", before, episode) and rank == 0:
            torch.save(model.state_dict(), os.path.join(args.output_dir, "rl_model_%d.pth" % episode))
    explorer.vec, explorer.device_memory, explorer.vec_stats = vec, dmem, stats
    return episode


def stride_cases(env, world):
    """data parallel: rank r plays cases r, r + world, ... (the reference's counter advances by one)"""
    if world <= 1:
        return
    orig = env.reset

    def reset(phase="test", test_case=None):
        ob = orig(phase, test_case)
        if env.case_counter[phase] >= 0:
            env.case_counter[phase] = (env.case_counter[phase] + world - 1) % env.case_size[phase]
        return ob
    env.reset = reset


if __name__ == "__main__":
    main()
