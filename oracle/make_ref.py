#!/usr/bin/env python
"""Recipe for oracle/_ref/: a verbatim, UNCOMMITTED copy of the reference's Python packages (crowd_nav/, crowd_sim/).

oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored, so the copy
travels to the GPU box, where /root/reference does not exist.  It serves ONE purpose there: bench.py's cpu_baseline times
the as-shipped Python reference (oracle/ref_bench.py, under oracle/ref_shim.py) on the box's host cores next to the C port.
    python oracle/make_ref.py            (in the build container; __graft_entry__.build() calls it when /root/reference exists)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("AEM_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def main():
    if not os.path.isdir(SRC):
        print("reference not present at %s: nothing to do" % SRC)
        return 1
    keep = (".py", ".config")
    for pkg in ("crowd_nav", "crowd_sim"):
        for root, dirs, files in os.walk(os.path.join(SRC, pkg)):
            dirs[:] = [d for d in dirs if d not in ("__pycache__", "data", "attachments")]
            for f in files:
                if f.endswith(keep):
                    rel = os.path.relpath(os.path.join(root, f), SRC)
                    out = os.path.join(DST, rel)
                    os.makedirs(os.path.dirname(out), exist_ok=True)
                    shutil.copyfile(os.path.join(root, f), out)
    print("copied the reference's crowd_nav / crowd_sim sources to %s (git-ignored)" % DST)
    return 0


if __name__ == "__main__":
    sys.exit(main())
