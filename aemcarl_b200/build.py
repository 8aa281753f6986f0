"""Build libaemcarl_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m aemcarl_b200.build [--force] [--verbose]

The shared library is git-ignored but travels with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libaemcarl_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + ARCH
# orca.cu: every FP32 op individually rounded (bit parity with the oracle); see the file header
SOURCES = {
    "lookahead.cu": [],
    "lookahead_tc.cu": [],
    "lookahead_h16.cu": [],
    "lookahead_q16.cu": [],
    "train.cu": ["-DAEM_TRAIN_TIMELINE"] if os.environ.get("AEM_TRAIN_TIMELINE") else [],
    "scenario.cu": ["-fmad=false"],   # Python float arithmetic of the scenario draw, op by op
    "env.cu": ["-fmad=false"],   # FP64 env arithmetic is a sequence of individually rounded Python float ops
    "orca.cu": ["-fmad=false"],
    "capi.cu": [],
}


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, "aem_common.cuh"), os.path.join(CSRC, "tc_common.cuh"), os.path.join(HERE, "..", "include", "aemcarl_b200.h"),
               os.path.abspath(__file__)]
    objs = []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc()] + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if force or _stale(LIB, objs):
        cmd = [nvcc(), "-shared", "-o", LIB] + objs + ARCH + ["-cudart", "static"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
