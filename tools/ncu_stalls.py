"""Whole-kernel stall-reason breakdown from an ncu report (source page, all samples)."""
import csv, io, subprocess, sys, collections
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], check=True, capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
tot = collections.Counter(); inst = 0
for r in rd:
    for k, v in r.items():
        if k and k.startswith("stall_") and "Not Issued" not in k and v: tot[k] += int(v)
    inst += int(r["Instructions Executed"] or 0)
s = sum(tot.values())
print("warp instructions", inst, "samples", s)
for k, v in tot.most_common(): print("  %-24s %5.1f%%" % (k, 100.0 * v / s))
