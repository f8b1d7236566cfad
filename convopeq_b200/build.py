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


def build_knobs() -> list[str]:
    """-D flags from the environment for the kernels' compile-time tuning knobs (the macros the sources guard with #ifndef
    CPQ_...); run-time knobs of the same prefix (read with getenv by the library) are not compile flags."""
    import re
    known = set()
    for d in DEPS:
        with open(d) as f:
            known.update(re.findall(r"^#ifndef (CPQ_[A-Z0-9_]+)", f.read(), flags=re.M))
    return [f"-D{k}={v}" for k, v in sorted(os.environ.items()) if k in known and v.isdigit()]


def source_hash(extra: list[str]) -> str:
    import hashlib
    h = hashlib.sha256()
    for d in sorted(os.path.abspath(x) for x in DEPS):
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + extra).encode())
    return h.hexdigest()


def up_to_date(extra: list[str] | None = None) -> bool:
    """libcpq.so exists and was built from exactly these sources and flags (a content hash beside it: file times do not survive
    every way a tree gets copied, and a stale-looking library must not send eight ranks into nvcc at once)."""
    if not os.path.exists(LIB):
        return False
    if not os.path.exists(LIB + ".hash"):   # a library without its hash file (copied on its own): fall back on the file times
        t = os.path.getmtime(LIB)
        return all(os.path.getmtime(d) <= t for d in DEPS)
    with open(LIB + ".hash") as f:
        return f.read().strip() == source_hash(build_knobs() if extra is None else extra)


def build(force: bool = False, verbose: bool = False, out: str | None = None, defines: dict | None = None) -> str:
    """Build libcpq.so.  `out` + `defines` build a tuning variant elsewhere (selected at run time with CPQ_LIB).
    Safe under concurrent callers (one process per GPU importing the package at once): an exclusive lock around the check and
    the compile, and the library appears under its name only when it is complete."""
    import fcntl
    extra = build_knobs()
    if out is None and not force and up_to_date(extra):
        return LIB
    extra_all = extra + [f"-D{k}={v}" for k, v in (defines or {}).items()]
    target = out or LIB
    os.makedirs(os.path.dirname(target), exist_ok=True)
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if out is None and not force and up_to_date(extra):   # another process built it while this one waited
                return LIB
            tmp = f"{target}.tmp.{os.getpid()}"
            cmd = [nvcc_path()] + NVCC_FLAGS + extra_all + ["-o", tmp] + SOURCES
            res = subprocess.run(cmd, capture_output=True, text=True)
            log = res.stdout + res.stderr
            with open(os.path.join(HERE, "build.log") if target == LIB else target + ".log", "w") as f:   # build.log describes libcpq.so only
                f.write(" ".join(cmd).replace(tmp, target) + "\n" + log)
            if verbose or res.returncode != 0:
                print(log[-8000:])
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libcpq.so")
            os.replace(tmp, target)
            if target == LIB:
                with open(LIB + ".hash.tmp", "w") as f:
                    f.write(source_hash(extra) + "\n")
                os.replace(LIB + ".hash.tmp", LIB + ".hash")
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return target


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
