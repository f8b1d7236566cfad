"""Deterministic synthetic inputs shared by the oracle, golden and GPU tests."""
from __future__ import annotations

import numpy as np

# EQProcessor::DEFAULT_FREQS (eqprocessor/EQProcessor.h:158-163)
DEFAULT_FREQS = [25.0, 40.0, 63.0, 100.0, 160.0, 250.0, 400.0, 630.0, 1000.0, 1600.0,
                 2500.0, 4000.0, 6300.0, 10000.0, 11000.0, 12500.0, 14000.0, 16500.0, 18000.0, 19500.0]


def noise(n: int, seed: int, amp: float = 0.1) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal(n) * amp


def log_sweep(n: int, sr: float, f0: float = 20.0, f1: float = 20000.0, amp: float = 0.5):
    t = np.arange(n) / sr
    dur = n / sr
    k = np.log(f1 / f0)
    ph = 2 * np.pi * f0 * dur / k * (np.exp(t / dur * k) - 1.0)
    return amp * np.sin(ph), amp * np.cos(ph)


def impulse(n: int, at: int = 0, amp: float = 1.0) -> np.ndarray:
    x = np.zeros(n)
    x[at] = amp
    return x


def silence_then_step(n: int, at: int, amp: float = 0.25) -> np.ndarray:
    x = np.zeros(n)
    x[at:] = amp
    return x


def synth_ir(n: int, seed: int) -> np.ndarray:
    """N(0,1)/sqrt(len) with an exponential decay, tau = len/6."""
    g = np.random.default_rng(seed)
    return g.standard_normal(n) / np.sqrt(n) * np.exp(-np.arange(n) / (n / 6.0))


def band_params(seed: int, stress: bool = False, modes=None, types=None, enabled=None, flat=None):
    """20 bands at DEFAULT_FREQS: band 0 LowShelf, 1..18 Peaking, 19 HighShelf; gains U(-6,6) dB, Q U(0.5,4).
    stress: Q = 20, +-24 dB alternating. Returns list of dicts."""
    g = np.random.default_rng(seed)
    out = []
    for i in range(20):
        t = 0 if i == 0 else (2 if i == 19 else 1)
        if types is not None:
            t = types[i]
        gain = float(g.uniform(-6, 6))
        q = float(g.uniform(0.5, 4))
        if stress:
            gain = 24.0 if i % 2 == 0 else -24.0
            q = 20.0
        if flat is not None and i in flat:
            gain = 0.004 * (1 if i % 2 else -1)   # inside createBandNode's 0.01 dB skip window (node path only)
        out.append(dict(frequency=DEFAULT_FREQS[i], gain=gain, q=q, enabled=True if enabled is None else bool(enabled[i]),
                        type=t, channel_mode=0 if modes is None else modes[i]))
    return out


def to_eqband(params):
    from oracle.bindings import EqBand
    return [EqBand(p["frequency"], p["gain"], p["q"], int(p["enabled"]), p["type"], p["channel_mode"]) for p in params]


def to_band(params):
    from convopeq_b200.engine import Band
    return [Band(p["frequency"], p["gain"], p["q"], p["enabled"], p["type"], p["channel_mode"]) for p in params]
