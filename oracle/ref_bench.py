#!/usr/bin/env python
"""Throughput of the AS-SHIPPED Python reference on host cores (SURVEY 8d "CPU baseline"):
N independent single-thread processes, each looping `robot.act(ob)` + `env.step(action)` of the UNMODIFIED reference
(MultiHumanRL.predict: 81 x {onestep_lookahead, rotate, batch-1 forward}; CrowdSim.step) over disjoint test cases, with
oracle/aem_oracle.c's ORCA standing in for the absent rvo2 wheel (oracle/ref_shim.py).  Test infrastructure only.

    AEM_REFERENCE=oracle/_ref python oracle/ref_bench.py --procs 16 --seconds 15     -> one JSON line
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def worker(rank, procs, seconds, humans, q):
    try:
        q.put(measure(rank, procs, seconds, humans))
    except BaseException as e:           # a dead worker must not leave the parent waiting on the queue
        q.put(("error", "%s: %s" % (type(e).__name__, e)))


def measure(rank, procs, seconds, humans):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    sys.path.insert(0, ROOT)
    import logging
    logging.disable(logging.CRITICAL)
    import torch
    torch.set_num_threads(1)
    from oracle import ref_shim
    policy = ref_shim.make_policy(seed=0)
    env, robot, cfg = ref_shim.make_env(human_num=humans)
    robot.set_policy(policy)
    policy.set_env(env)
    steps, case = 0, rank
    ob = env.reset("test", case)
    robot.act(ob)                                   # warm-up: numba compilation of the gridmap kernels
    ob = env.reset("test", case)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        action = robot.act(ob)
        ob, reward, done, info = env.step(action)
        steps += 1
        if done:
            case += procs
            ob = env.reset("test", case % 500)
    return steps, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--seconds", type=float, default=15.0)
    ap.add_argument("--humans", type=int, default=5)
    a = ap.parse_args()
    ref = os.environ.get("AEM_REFERENCE") or ("/root/reference" if os.path.isdir("/root/reference") else os.path.join(HERE, "_ref"))
    os.environ["AEM_REFERENCE"] = os.path.abspath(ref)
    # this is the CPU baseline: hide the GPU from the workers, so that the reference's hard-coded .cuda() calls are the
    # identity of oracle/ref_shim.py on a GPU box too (with a visible GPU they would move the inputs of a CPU model there)
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=worker, args=(r, a.procs, a.seconds, a.humans, q)) for r in range(a.procs)]
    t0 = time.perf_counter()
    for p in ps:
        p.start()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    bad = [r[1] for r in res if r[0] == "error"]
    if bad:
        print(json.dumps({"unavailable": "reference worker failed: " + bad[0][:300]}))
        return
    steps = sum(s for s, _ in res)
    rate = sum(s / t for s, t in res)
    print(json.dumps({"value": rate, "unit": "env-steps/s", "cores": a.procs, "kind": "reference",
                      "env_steps": steps, "per_core": rate / a.procs, "wall_s": time.perf_counter() - t0,
                      "sample": "the unmodified Python reference (MultiHumanRL.predict + CrowdSim.step, CPU, eval mode, "
                                "%d humans) in %d single-thread processes for %.0f s each; ORCA by oracle/aem_oracle.c"
                                % (a.humans, a.procs, a.seconds)}))


if __name__ == "__main__":
    main()
