"""Build libcodeps_photo.so in-tree with nvcc for sm_100a (B200).

    python -m codeps_b200.build [--force]

The shared library is a plain C-ABI library (include/codeps_photo.h); it links the CUDA runtime
statically and has no Python or torch dependency.  It is git-ignored but travels with the
working tree to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libcodeps_photo.so")
SOURCES = [os.path.join(CSRC, "cdp_api.cu")]
HEADERS = [os.path.join(CSRC, n) for n in ("cdp_common.h", "cdp_math.h", "cdp_kernels.h", "cdp_photo_tile.h", "cdp_plan.h", "cdp_flow.h", "cdp_c2c.h", "cdp_metrics.h")] + \
          [os.path.join(REPO, "include", "codeps_photo.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def nvcc_path() -> str:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS)


def build_native(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    """Compile the CUDA extension if it is missing or older than its sources."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra_flags, "-I", os.path.join(REPO, "include"), "-I", CSRC,
           "-o", LIB_PATH, *SOURCES]
    if verbose:
        print(" ".join(cmd), flush=True)
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    if verbose and (proc.stdout or proc.stderr):
        print(proc.stdout, proc.stderr, flush=True)
    return LIB_PATH


if __name__ == "__main__":
    path = build_native(force="--force" in sys.argv, verbose=True,
                        extra_flags=("-Xptxas", "-v") if "--ptxas" in sys.argv else ())
    print(path)
