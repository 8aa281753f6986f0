"""Data-parallel training check (BASELINE configs[4] shape, shortened): run under torchrun, every rank trains with the fused
optimisation step and the vectorised explorer; at the end the ranks compare their parameters (they must be bit-identical: same
initial broadcast, same all-reduced gradients) and rank 0 prints timings.
usage: torchrun --nproc-per-node N tools/dp_train_check.py [--vectorised 256] [--train_episodes 1024] [--updates_per_wave 400]"""
import argparse, json, os, sys, tempfile, time
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from aemcarl_b200.crowd_nav import train
ap = argparse.ArgumentParser()
ap.add_argument("--vectorised", type=int, default=256)
ap.add_argument("--train_episodes", type=int, default=1024)
ap.add_argument("--updates_per_wave", type=int, default=400)
ap.add_argument("--il_episodes", type=int, default=128)
ap.add_argument("--mode", default="fused", choices=["fused", "graph", "eager"])
a = ap.parse_args()
rank = int(os.environ.get("RANK", "0"))
out = tempfile.mkdtemp(prefix="aem_dp_%d_" % rank)
argv = ["--no_final_test", "--output_dir", out, "--il_episodes", str(a.il_episodes), "--il_epochs", "2", "--train_episodes", str(a.train_episodes),
        "--vectorised", str(a.vectorised), "--updates_per_wave", str(a.updates_per_wave)]
argv += {"fused": ["--fused_step"], "graph": ["--cuda_graph"], "eager": []}[a.mode]
torch.cuda.synchronize() if torch.cuda.is_initialized() else None
t0 = time.time()
res = train.main(argv)
torch.cuda.synchronize()
dt = time.time() - t0
flat = torch.from_numpy(res["model"].flat_weights()).cuda()
world = dist.get_world_size() if dist.is_initialized() else 1
same = True
if world > 1:
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
if rank == 0:
    waves = (a.train_episodes + a.vectorised * world - 1) // (a.vectorised * world)
    print(json.dumps({"world": world, "mode": a.mode, "seconds": dt, "waves": waves, "episodes": waves * a.vectorised * world,
                      "optimizer_steps": waves * a.updates_per_wave, "replicas_identical": bool(same),
                      "finite": bool(torch.isfinite(flat).all()), "memory": res["memory"]}))
assert same
