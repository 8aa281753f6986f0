#!/usr/bin/env python
"""Report of a lookahead_h16 timeline (libaemcarl_b200_tl.so, AEM_H16_TIMELINE=<file>): clock64 stamps of thread 0
(issuer warp), thread 32 (producer warp) and thread 255 of CTA 0 over one steady-state tile.
usage: python tools/h16_timeline_report.py gpurun_out/tl0.txt [source.cu]"""
import collections
import sys

rows = collections.defaultdict(list)
for l in open(sys.argv[1]):
    s, line, clk = l.split()
    rows[int(s)].append((int(line), int(clk)))
src = open(sys.argv[2] if len(sys.argv) > 2 else "aemcarl_b200/csrc/lookahead_h16.cu").read().split("\n")
names = {0: "thread 0 (issuer warp)", 1: "thread 32 (producer warp)", 2: "thread 255"}
for s in sorted(rows):
    r = rows[s]
    t0 = r[0][1]
    print("== %s: %d stamps, tile = %d cycles" % (names[s], len(r), r[-1][1] - t0))
    agg = collections.OrderedDict()
    for (l0, c0), (l1, c1) in zip(r[:-1], r[1:]):
        key = (l0, l1)
        a = agg.setdefault(key, [0, 0])
        a[0] += c1 - c0
        a[1] += 1
    tot = r[-1][1] - t0
    for (l0, l1), (cyc, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("  %5d -> %5d  x%-3d %7d cycles (%4.1f %%)  avg %6d   %s" % (l0, l1, n, cyc, 100.0 * cyc / tot, cyc // n,
                                                                        src[l0 - 1].strip()[:70]))
