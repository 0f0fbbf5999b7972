"""Build libadb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

`build(debug=True)` additionally builds libadb200_dbg.so with -DADB_DEBUG: the same code plus the in-kernel cycle accounting
and the timing-only experiment switches (ADB_DEBUG_FLAGS). Only tools/ load it (ADB_LIB=debug); the product library has no
run-time debug predicate.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "adb200.cu")
OUT = os.path.join(HERE, "libadb200.so")
OUT_DBG = os.path.join(HERE, "libadb200_dbg.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + \
       [os.path.join(os.path.dirname(HERE), "include", "adb200.h")]


def up_to_date(out=OUT):
    if not os.path.exists(out):
        return False
    t = os.path.getmtime(out)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force=False, verbose=False, debug=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    for out, extra in ((OUT, []),) + (((OUT_DBG, ["-DADB_DEBUG"]),) if debug else ()):
        if not force and up_to_date(out):
            continue
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-shared", "-Xcompiler", "-fPIC", *extra, "-o", out, SRC]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
