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
HEADERS = ["ptx.cuh", "gemm_tc.cuh", "attn_tc.cuh", "elementwise.cuh", "encoder.cuh", "d0.cuh", "prepare.cuh", "postprocess.cuh", "rk_tc.cuh", "sk_tc.cuh"]


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


def build(force: bool = False, verbose: bool = False, wait_log: "int | None" = None, out: "Path | None" = None) -> Path:
    """wait_log: record level of the bounded barrier waits (csrc/ptx.cuh; default 1, 2 = one record per stuck waiter);
    out: library path (default: the in-tree libsyncfusion_b200.so the package loads)."""
    target = Path(out) if out else LIB
    if not force and out is None and wait_log is None and not needs_build():
        return LIB
    level = wait_log if wait_log is not None else (int(os.environ["SFB_WAIT_LOG"]) if os.environ.get("SFB_WAIT_LOG") else None)
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
           "-DSFB_NO_FAST_MATH", *(["-DSFB_RK_PIPE"] if os.environ.get("SFB_RK_PIPE") else []), *(["-DSFB_ATTN_TL"] if os.environ.get("SFB_ATTN_TL") else []),
           *([f"-DSFB_WAIT_LOG={level}"] if level is not None else []), "-Xcompiler", "-fPIC", "-shared",
           "-Xptxas", "-v" if verbose else "-O3", "-o", str(target)] + [str(CSRC / s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return target


if __name__ == "__main__":
    lvl = int(sys.argv[sys.argv.index("--wait-log") + 1]) if "--wait-log" in sys.argv else None
    dst = Path(sys.argv[sys.argv.index("-o") + 1]) if "-o" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, wait_log=lvl, out=dst))
