"""Builds ark_ec_vrfs_b200/libvrfs_b200.so (sm_100a only) in-tree with nvcc.  No GPU needed to build."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "vrfs_b200.cu")
LIB = os.path.join(HERE, "libvrfs_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "--cudart", "static"]


def sources():
    out = [os.path.join(HERE, "..", "include", "vrfs_b200.h")]
    for root, _, files in os.walk(os.path.join(HERE, "csrc")):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h"))]
    return out


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def build(force=False, verbose=False, extra=(), out=None):
    if out is None and not force and not stale():
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + list(extra) + ["-o", out or LIB, SRC]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    build(force="--force" in sys.argv, verbose=True, extra=[a for a in sys.argv[1:] if a.startswith("-D") or a.startswith("-X")],
          out=outs[0] if outs else None)
