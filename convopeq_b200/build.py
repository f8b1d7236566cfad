"""Build convopeq_b200/libcpq.so (the C-ABI library of include/cpq.h) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with the
repo snapshot.  `python -m convopeq_b200.build [--force] [--verbose]`
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcpq.so")
SOURCES = [os.path.join(CSRC, "cpq_engine.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "cpq_fft.cuh"), os.path.join(CSRC, "cpq_fft16.cuh"), os.path.join(CSRC, "cpq_fft_large.cuh"), os.path.join(CSRC, "cpq_mac.cuh"), os.path.join(CSRC, "cpq_eq.cuh"),
                  os.path.join(CSRC, "cpq_plan.hpp"),
                  os.path.join(HERE, "..", "include", "cpq.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++20",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
    "--fmad=true",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: convopeq_b200 needs the CUDA toolkit to build")
    return p


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False, out: str | None = None, defines: dict | None = None) -> str:
    """Build libcpq.so.  `out` + `defines` build a tuning variant elsewhere (selected at run time with CPQ_LIB)."""
    if out is None and not force and up_to_date():
        return LIB
    extra = [f"-D{k}={v}" for k, v in os.environ.items() if k.startswith("CPQ_") and k.isupper() and v.isdigit()]   # tuning knobs
    extra += [f"-D{k}={v}" for k, v in (defines or {}).items()]
    target = out or LIB
    os.makedirs(os.path.dirname(target), exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + ["-o", target] + SOURCES
    out_ = subprocess.run(cmd, capture_output=True, text=True)
    out = out_
    log = out.stdout + out.stderr
    with open(os.path.join(HERE, "build.log") if target == LIB else target + ".log", "w") as f:   # build.log describes libcpq.so only
        f.write(" ".join(cmd) + "\n" + log)
    if verbose or out.returncode != 0:
        print(log[-8000:])
    if out.returncode != 0:
        raise RuntimeError("nvcc failed building libcpq.so")
    return target


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
