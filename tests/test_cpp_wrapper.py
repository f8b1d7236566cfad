"""include/cpq.hpp (C++20 host wrapper) compiles with g++ -std=c++20 and behaves: fails loudly without a GPU,
and on a GPU produces the same numbers as the ctypes path."""
import os
import subprocess
import ctypes as C

import numpy as np
import pytest

from convopeq_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_wrapper_test")


def _build():
    capi.load()
    libdir = os.path.join(ROOT, "convopeq_b200")
    cmd = ["g++", "-std=c++20", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_wrapper_test.cpp"),
           "-o", EXE, "-L", libdir, "-l:libcpq.so", f"-Wl,-rpath,{libdir}"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]


def test_cpp_wrapper_compiles_and_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _build()
    out = subprocess.run([EXE, "0"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "no CPU fallback" in out.stdout


def _lcg(seed):
    seed[0] = (seed[0] * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
    return ((seed[0] >> 11) & ((1 << 53) - 1)) / float(1 << 53) - 0.5


@pytest.mark.gpu
def test_cpp_wrapper_matches_python_path(checker):
    from convopeq_b200.engine import ConvoPeqEngine, Band
    from oracle.bindings import FilterSpec as OFilterSpec, EqBand
    _build()
    out = subprocess.run([EXE, "1"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("OK gpu"), out.stdout + out.stderr
    fields = dict(kv.split("=") for kv in out.stdout.split(":")[1].split())
    T, B, ir_len = 8192, 512, 20000
    seed = [12345]
    irL, irR = np.zeros(ir_len), np.zeros(ir_len)
    for i in range(ir_len):
        irL[i] = _lcg(seed) * np.exp(-i / 3000.0) * 0.05
        irR[i] = _lcg(seed) * np.exp(-i / 3000.0) * 0.05
    l, r = np.zeros(T), np.zeros(T)
    for i in range(T):
        l[i] = 0.2 * _lcg(seed)
        r[i] = 0.2 * _lcg(seed)
    freqs = [20, 32, 50, 80, 125, 200, 315, 500, 800, 1250, 2000, 3150, 5000, 8000, 12500, 16000, 19000, 20000, 22000, 24000]
    bands = [EqBand(float(freqs[b]), 3.0 if b % 2 else -2.5, np.float32(1.0) + np.float32(0.1) * np.float32(b), 1, 0 if b == 0 else (2 if b == 19 else 1), 0)
             for b in range(20)]
    from oracle.bindings import Oracle
    want = checker.chain_run((irL, irR), bands, np.stack([l, r]), 48000.0, B, OFilterSpec(), total_gain_db=-1.0, makeup=1.1)
    s = float(np.sum(want[0] - want[1]))
    q = float(np.sum(want[0] ** 2 + want[1] ** 2))
    assert abs(float(fields["sum"]) - s) <= 1e-9 and abs(float(fields["sq"]) - q) <= 1e-9 * max(1.0, q)
    assert int(fields["latency"]) == 512
