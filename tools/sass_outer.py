"""Static SASS size per OUTERMOST source line of the kernel (inlined callees charged to their call site).
usage: python tools/sass_outer.py obj.o [top]"""
import collections, os, re, subprocess, sys, tempfile
obj = os.path.abspath(sys.argv[1]); top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, check=True, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], cwd=d, check=True, capture_output=True, text=True).stdout
cnt = collections.Counter(); last = None
for l in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: last = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l) and last: cnt[last] += 1
tot = sum(cnt.values()); print("total instructions", tot, "=", tot * 16 // 1024, "KB")
src = open(sys.argv[1].replace("_obj/", "").replace(".o", ".cu")).read().split("\n")
for k, v in cnt.most_common(top):
    text = src[k[1] - 1].strip()[:90] if k[0].endswith(".cu") and k[1] <= len(src) else ""
    print("%-18s %4d %5d (%4.1f KB)  %s" % (k[0], k[1], v, v * 16 / 1024, text))
