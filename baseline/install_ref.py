"""Make the UNMODIFIED reference importable where `/root/reference` does not exist (the GPU box).

The reference ships no build system (no setup.py / pyproject: `pip install /root/reference` has nothing to build), so
the "install" is a verbatim copy of the modules on the hot path into `baseline/_ref/` — git-ignored (never part of this
repository's history), not gpurun-ignored (it travels with the snapshot). Run by `__graft_entry__.build()` in the build
container; a no-op where the reference tree is absent. `bench.py --impl reference` and `tools/ref_gpu.py` import the
reference's own classes from there through `oracle/ref_loader.py` and fall back to the oracle port (saying so) without it.
"""
import os
import shutil

ROOT = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ADB_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "_ref")
DIRS = ["src/models/components", "src/models/backbones"]
FILES = ["src/__init__.py", "src/models/__init__.py", "src/models/phema.py", "LICENSE"]


def install(force=False):
    if not os.path.isdir(os.path.join(SRC, "src", "models", "components")):
        return None
    if os.path.isdir(DST) and not force:
        return DST
    shutil.rmtree(DST, ignore_errors=True)
    for d in DIRS:
        os.makedirs(os.path.join(DST, d), exist_ok=True)
        for f in os.listdir(os.path.join(SRC, d)):
            p = os.path.join(SRC, d, f)
            if os.path.isfile(p) and f.endswith(".py"):
                shutil.copyfile(p, os.path.join(DST, d, f))
    for f in FILES:
        p = os.path.join(SRC, f)
        if os.path.isfile(p):
            os.makedirs(os.path.dirname(os.path.join(DST, f)), exist_ok=True)
            shutil.copyfile(p, os.path.join(DST, f))
    return DST


if __name__ == "__main__":
    print(install(force=True))
