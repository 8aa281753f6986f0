"""SASS instruction count per source line of a kernel object (built with -lineinfo).
usage: python tools/sass_lines.py aemcarl_b200/csrc/_obj/lookahead_h16.o [top]"""
import collections, os, re, subprocess, sys, tempfile
obj = os.path.abspath(sys.argv[1]); top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=d, check=True, capture_output=True, text=True).stdout
cur = None; cnt = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for l in txt.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', l)
    if m and cur:
        cnt[cur] += 1; ops[cur][m.group(1).split('.')[0]] += 1
print("total", sum(cnt.values()))
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]:
    print("%-22s %5d  %5d  %s" % (k[0], k[1], v, " ".join("%s:%d" % o for o in ops[k].most_common(6))))
