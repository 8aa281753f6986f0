"""User-facing entry points end to end: train (fused, vectorised) -> checkpoint -> test (serial and vectorised, options)."""
import sys, tempfile, os
sys.path.insert(0, ".")
from aemcarl_b200.crowd_nav import train, test
out = tempfile.mkdtemp()
res = train.main(["--no_final_test", "--output_dir", out, "--il_episodes", "32", "--il_epochs", "2", "--train_episodes", "128", "--vectorised", "64",
                  "--updates_per_wave", "20", "--fused_step", "--cuda_graph"])
ckpt = os.path.join(out, "rl_model.pth")
a = test.main(["--model_weights", ckpt, "--cases", "3"])
b = test.main(["--model_weights", ckpt, "--vectorised", "--cases", "100", "--precision", "fp16x3"])
c = test.main(["--model_weights", ckpt, "--vectorised", "--cases", "64", "--square", "--human_num", "10"])
d = test.main(["--model_weights", ckpt, "--vectorised", "--cases", "64", "--visible"])
print("OK", b["success"], b["collision"], b["timeout"], c["success"], d["success"])
