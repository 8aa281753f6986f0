"""CPU, world_size 2, gloo: the one collective of the system (flat gradient all-reduce between backward and
optimizer.step, trainer.py:55-56,79-80) keeps replicas bit-identical; env sharding needs no exchange."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aemcarl_b200.crowd_nav.default_configs import policy_config
    from aemcarl_b200.crowd_nav.policy.policy_factory import policy_factory
    from aemcarl_b200.crowd_nav.utils.memory import ReplayMemory
    from aemcarl_b200.crowd_nav.utils.trainer import Trainer, broadcast_parameters
    from aemcarl_b200 import scenarios
    torch.manual_seed(100 + rank)                       # different init per rank ...
    policy = policy_factory["actenvcarl"]()
    policy.configure(policy_config())
    model = policy.get_model()
    broadcast_parameters(model, 0)                      # ... made identical by the start-up broadcast
    mem = ReplayMemory(1000)
    g = torch.Generator().manual_seed(rank)             # rank-local replay data (env shards are disjoint)
    for _ in range(40):
        mem.push((torch.randn(5, 13, generator=g), torch.randn(1, generator=g)))
    trainer = Trainer(model, mem, torch.device("cpu"), 10)
    trainer.set_learning_rate(0.01, "Adam")
    torch.manual_seed(7 + rank)
    trainer.optimize_batch(3)
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    # disjoint scenario shards: rank r owns cases r*B .. r*B+B-1
    rb, hs = scenarios.generate_pool("train", rank * 4, 4, 5, "circle_crossing")
    if rank == 0:
        out.put((bool(torch.equal(gathered[0], gathered[1])), float((gathered[0] - flat).abs().max())))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gradient_allreduce_keeps_replicas_identical():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    same, diff = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same and diff == 0.0
