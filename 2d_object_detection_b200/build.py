"""Build libfrcnn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Flags that matter for parity: ``-fmad=false`` (no FMA contraction: every fp32 op rounds separately,
like TF's CPU kernels -- SURVEY traps T5/T6) and no ``--use_fast_math``.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libfrcnn_b200.so")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
    "-shared", "--threads", "0",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libfrcnn_b200.so cannot be built (no CPU fallback exists)")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _defines() -> str:
    return " ".join(sorted(os.environ.get("FRCNN_DEFINES", "").split()))


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    try:  # a library built with other -D switches (debug / experiment builds) is stale
        with open(LIB_PATH + ".defines") as f:
            if f.read().strip() != _defines():
                return True
    except OSError:
        if _defines():
            return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-o", LIB_PATH, *sources()]
    for d in os.environ.get("FRCNN_DEFINES", "").split():  # e.g. FRCNN_DEFINES=FRCNN_DEBUG_SPIN (bounded mbarrier waits)
        cmd[1:1] = ["-D" + d]
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stderr[-4000:])
    with open(LIB_PATH + ".defines", "w") as f:
        f.write(_defines() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
