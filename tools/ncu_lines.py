"""Aggregate an ncu source-page export (SASS level) per line of the KERNEL body: the outermost source line of every
instruction (through the inlined-at chain of nvdisasm -gi) -> executed warp instructions, stall samples, top stalls.
usage: python tools/ncu_lines.py <report.ncu-rep> <object.o> [top] [byline|-] [kernel-name substring when the object holds several]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj = sys.argv[1], os.path.abspath(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], cwd=d, check=True, capture_output=True, text=True).stdout
addr2line = {}; last = None
want = sys.argv[5] if len(sys.argv) > 5 else None; insec = want is None
for l in dis.splitlines():
    if l.startswith("//---------------------") and want is not None:
        insec = want in l
        continue
    if not insec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        last = (os.path.basename(m.group(1)), int(m.group(2)))       # the last annotation before an instruction is the outermost
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
    if m and last: addr2line[int(m.group(1), 16)] = last
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], check=True, capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
stall_cols = None
agg = collections.defaultdict(lambda: collections.Counter())
base = None
for r in rd:
    if stall_cols is None:
        stall_cols = [c for c in r.keys() if c and c.startswith("stall_") and "Not Issued" not in c]
    try: a = int(r["Address"], 16) if r["Address"].startswith("0x") else int(r["Address"])
    except Exception: continue
    if base is None: base = a
    key = addr2line.get(a - base, ("?", 0))
    c = agg[key]
    c["inst"] += int(r["Instructions Executed"] or 0)
    c["samples"] += int(r["# Samples"] or 0)
    for s in stall_cols: c[s] += int(r[s] or 0)
tot_i = sum(c["inst"] for c in agg.values()); tot_s = sum(c["samples"] for c in agg.values())
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
rows = sorted(agg.items(), key=lambda kv: -kv[1]["samples"])
if len(sys.argv) > 4 and sys.argv[4] == "byline": rows = sorted(agg.items())
for k, c in rows[:top]:
    st = sorted(((c[s], s[6:]) for s in stall_cols), reverse=True)[:3]
    print("%-18s %4d  inst %5.2f%%  samples %5.2f%%  %s" % (k[0], k[1], 100.0 * c["inst"] / tot_i, 100.0 * c["samples"] / tot_s,
          " ".join("%s:%.0f%%" % (n, 100.0 * v / max(c["samples"], 1)) for v, n in st)))
