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


def build_variant(name, defines, verbose=False):
    """A debug / experiment variant of the library (e.g. the timeline build of lookahead_h16.cu):
    tools/bin/libaemcarl_b200_<name>.so, selected at run time with AEM_LIB_PATH (aemcarl_b200/_lib.py)."""
    obj = os.path.join(CSRC, "_obj_" + name)
    out_dir = os.path.join(HERE, "..", "tools", "bin")
    os.makedirs(obj, exist_ok=True)
    os.makedirs(out_dir, exist_ok=True)
    objs = []
    for src, extra in SOURCES.items():
        o = os.path.join(obj, src.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc()] + COMMON + extra + ["-D" + d for d in defines] + ["-c", os.path.join(CSRC, src), "-o", o]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    lib = os.path.abspath(os.path.join(out_dir, "libaemcarl_b200_%s.so" % name))
    subprocess.check_call([nvcc(), "-shared", "-o", lib] + objs + ARCH + ["-cudart", "static"])
    return lib


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, "aem_common.cuh"), os.path.join(CSRC, "tc_common.cuh"), os.path.join(CSRC, "orca_device.cuh"),
               os.path.join(HERE, "..", "include", "aemcarl_b200.h"),
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
    if "--variant" in sys.argv:          # python -m aemcarl_b200.build --variant tl AEM_H16_TIMELINE
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:], verbose="--verbose" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
