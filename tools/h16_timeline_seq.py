#!/usr/bin/env python
"""Sequential view of a lookahead_h16 timeline: for one thread slot, every interval between two MARK() stamps in order.
usage: python tools/h16_timeline_seq.py file slot [source.cu]"""
import sys
rows = [l.split() for l in open(sys.argv[1])]
slot = int(sys.argv[2])
src = open(sys.argv[3] if len(sys.argv) > 3 else "aemcarl_b200/csrc/lookahead_h16.cu").read().split("\n")
r = [(int(l), int(c)) for s, l, c in rows if int(s) == slot]
t0 = r[0][1]
prev = None
kinds = {"barrier+issue": 0, "mma wait": 0, "produce": 0, "epilogue": 0}
for i, (l, c) in enumerate(r):
    txt = src[l - 1].strip()
    if prev is not None:
        pl, pc = prev
        ptxt = src[pl - 1].strip()
        d = c - pc
        if "WAIT_EVENT" in ptxt and pl == l:
            kind = "mma wait"
        elif "PRODUCE" in ptxt and "WAIT_EVENT" in txt:
            kind = "produce"
        elif "PRODUCE" in txt:
            kind = "barrier+issue"
        else:
            kind = "epilogue"
        kinds[kind] += d
        print("%7d  +%6d  %-14s %4d -> %4d   %s" % (pc - t0, d, kind, pl, l, txt[:60]))
    prev = (l, c)
tot = r[-1][1] - t0
print("total %d cycles: " % tot + ", ".join("%s %d (%.0f %%)" % (k, v, 100.0 * v / tot) for k, v in kinds.items()))
