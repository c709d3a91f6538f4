"""Compile csrc/ into libpgibbs.so for sm_100a (in-tree, so the .so travels with the repo snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("PGIBBS_LIB_OUT") or os.path.join(HERE, "libpgibbs.so")   # (override: A/B builds for tooling)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + \
        [os.path.join(HERE, "..", "include", "pgibbs.h")]


def is_stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in sources())


def build_library(force=False, verbose=False):
    if not force and not is_stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("PGIBBS_NVCC_EXTRA", "").split()   # e.g. -DPGIBBS_FA_POLY_EVERY=3 for kernel A/B builds
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", OUT, os.path.join(CSRC, "engine.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libpgibbs.so")
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose=True)
    print("built", OUT)
