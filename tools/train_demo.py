"""A short full training run (imitation warm-up + vectorised RL with the fused optimisation step) and the 500 seeded test cases
before / after, with wall-clock times -- the shape of BASELINE configs[4] on one GPU.
usage: python tools/train_demo.py [--train_episodes 16384] [--vectorised 2048] [--updates_per_wave 2000]"""
import argparse, json, os, sys, tempfile, time
import torch
sys.path.insert(0, ".")
from aemcarl_b200.crowd_nav import train, test
ap = argparse.ArgumentParser()
ap.add_argument("--il_episodes", type=int, default=2048)
ap.add_argument("--il_epochs", type=int, default=20)
ap.add_argument("--train_episodes", type=int, default=16384)
ap.add_argument("--vectorised", type=int, default=2048)
ap.add_argument("--updates_per_wave", type=int, default=2000)
ap.add_argument("--save", default=None, help="write the trained flat weights (.npy) here, e.g. for bench.py --weights")
ap.add_argument("--dropout", action="store_true", help="optimise in train() mode: the reference's regime (dropout p = 0.1 live)")
ap.add_argument("--target_dropout", action="store_true", help="the target network in train() mode as well (explorer.py:132-139)")
a = ap.parse_args()
out = tempfile.mkdtemp()
t0 = time.time()
res = train.main((["--dropout"] if a.dropout else []) + (["--target_dropout"] if a.target_dropout else []) + ["--no_final_test", "--output_dir", out, "--il_episodes", str(a.il_episodes), "--il_epochs", str(a.il_epochs), "--train_episodes",
                  str(a.train_episodes), "--vectorised", str(a.vectorised), "--updates_per_wave", str(a.updates_per_wave), "--fused_step"])
torch.cuda.synchronize()
t_train = time.time() - t0
if a.save:
    import numpy as np
    np.save(a.save, res["model"].flat_weights())
rows = {}
for name, ckpt in (("imitation", "il_model.pth"), ("final", "rl_model.pth")):
    t0 = time.time()
    r = test.main(["--model_weights", os.path.join(out, ckpt), "--vectorised", "--cases", "500", "--precision", "fp16x3"])
    rows[name] = {"success": r["success"] / 500, "collision": r["collision"] / 500, "timeout": r["timeout"] / 500,
                  "seconds": time.time() - t0}
# the trained weights through every arithmetic variant of the lookahead kernel: per-case outcomes of the 500 test cases
codes = {}
for prec in ("fp32", "tf32x3", "fp16x3"):
    codes[prec] = test.main(["--model_weights", os.path.join(out, "rl_model.pth"), "--vectorised", "--cases", "500", "--precision", prec])["codes"]
rows["outcomes_identical_across_precisions"] = {p: int((codes[p] != codes["fp32"]).sum()) for p in codes}
print(json.dumps({"dropout_in_training": bool(a.dropout), "dropout_in_target_network": bool(a.target_dropout), "train_seconds": t_train, "il_episodes": a.il_episodes, "rl_episodes": a.train_episodes,
                  "optimizer_steps_rl": (a.train_episodes // a.vectorised) * a.updates_per_wave, "memory": res["memory"], "test": rows}))
