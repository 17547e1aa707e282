"""Builds libsyncfusion_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libsyncfusion_b200.so"
SOURCES = ["sfb.cu"]
HEADERS = ["ptx.cuh", "gemm_tc.cuh", "attn_tc.cuh", "elementwise.cuh", "d0.cuh", "prepare.cuh", "rk_tc.cuh", "sk_tc.cuh"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES + HEADERS] + [HERE.parent / "include" / "syncfusion_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
           "--use_fast_math" if False else "-DSFB_NO_FAST_MATH", *(["-DSFB_RK_PIPE"] if os.environ.get("SFB_RK_PIPE") else []), "-Xcompiler", "-fPIC", "-shared",
           "-Xptxas", "-v" if verbose else "-O3", "-o", str(LIB)] + [str(CSRC / s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
