#!/usr/bin/env python
"""Build libdet_b200.so (sm_100a only) in-tree with nvcc.  No torch headers, no JIT cache: the .so sits next to the
Python package so it travels to the GPU box with the repository snapshot."""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "det_b200", "_lib")
OUT = os.path.join(OUT_DIR, "libdet_b200.so")
STAMP = OUT + ".srchash"  # hash of the sources + flags the .so was built from (mtimes do not survive a snapshot copy)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",              # IEEE operation order == the reference's CPU torch arithmetic
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "--shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "--threads", "0",           # the translation units are compiled in parallel
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def dependencies():
    return sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
        sorted(glob.glob(os.path.join(HERE, "..", "include", "*.h")))


def source_hash():
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in dependencies():
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build():
    """True unless libdet_b200.so exists and was built from exactly the sources and flags in the tree now."""
    if not os.path.exists(OUT) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False, debug_phases=False):
    """debug_phases=True builds libdet_b200_dbg.so with the DET_MARK phase timeline compiled in (profiling only)."""
    out = OUT.replace(".so", "_dbg.so") if debug_phases else OUT
    if not debug_phases and not force and not needs_build():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = (["-Xptxas", "-v"] if verbose else []) + (["-DDET_DEBUG_PHASES"] if debug_phases else [])
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", out] + sources()
    print("[det_b200] " + " ".join(cmd), file=sys.stderr)
    if not debug_phases and os.path.exists(STAMP):
        os.remove(STAMP)
    subprocess.check_call(cmd)
    if not debug_phases:
        with open(STAMP, "w") as f:
            f.write(source_hash() + "\n")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug_phases="--debug-phases" in sys.argv))
