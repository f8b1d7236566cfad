"""The C restatement (oracle/cpq_oracle.c) against golden vectors produced by the reference's own code
(tests/golden/make_golden.py).  Runs everywhere, including where /root/reference does not exist."""
import os

import numpy as np
import pytest

from oracle.bindings import FilterSpec
from tests import signals
from tests.golden.cases import (CONV_CASES, EQ_CASES, CHAIN_CASES, OUTPUT_CASES, FULL_CHAIN_CASES, conv_inputs, eq_inputs, chain_inputs,
                                output_inputs, DITHER_CASES, dither_inputs)

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden.npz"))
TOL = 1e-12   # restatement vs reference: same algorithm, different FFT rounding (SURVEY 8c measured 5.8e-15)


@pytest.mark.parametrize("name", sorted(CONV_CASES))
def test_convolver_restatement_matches_golden(oracle, name):
    c = CONV_CASES[name]
    ir, x = conv_inputs(c)
    spec = FilterSpec(**c["spec"]) if c["spec"] is not None else None
    y, lay = oracle.nuc_run(ir, x, c["block"], scale=c.get("scale", 1.0), spec=spec, direct_head=c.get("direct_head", False))
    assert np.abs(y - GOLD["conv/" + name]).max() <= TOL
    got = np.array([[l["part_size"], l["num_parts_ir"], l["parts_per_callback"], l["output_delay_samples"]] for l in lay["layers"]])
    assert np.array_equal(got, GOLD["conv_layout/" + name])
    assert np.array_equal(np.array(lay["gains"]), GOLD["conv_gains/" + name])


def test_impulse_onset_and_tail_alignment_in_golden():
    """The reference's own output: L0 speaks at n = 511 (zero latency), the tail first contributes at 511 + D1."""
    y = GOLD["conv/impulse_at_511"]
    assert np.flatnonzero(np.abs(y) > 1e-12)[0] == 511   # FFT rounding leaves ~1e-18 before the onset
    ir, _ = conv_inputs(CONV_CASES["impulse_at_511"])
    # h_eff = h[0:5760] (+) 1.4375 * h[5760:] shifted to D1 = 7168
    g1 = float(GOLD["conv_gains/impulse_at_511"][1])
    heff = np.zeros(16384)
    heff[511:511 + 5760] += ir[:5760]
    n = 16384 - (511 + 7168)
    heff[511 + 7168:] += g1 * ir[5760:5760 + n]
    assert np.abs(y - heff).max() < 1e-13


@pytest.mark.parametrize("name", sorted(EQ_CASES))
def test_eq_restatement_matches_golden(oracle, name):
    c = EQ_CASES[name]
    bands, xl, xr = eq_inputs(c)
    l, r, st = oracle.eq_run(signals.to_eqband(bands), xl, xr, c["sr"], c["block"], **c.get("kw", {}))
    g = GOLD["eq/" + name]
    assert np.abs(l - g[0]).max() <= 1e-12 * max(1.0, np.abs(g).max()) and np.abs(r - g[1]).max() <= 1e-12 * max(1.0, np.abs(g).max())
    assert np.abs(st - GOLD["eq_state/" + name]).max() <= 1e-11


@pytest.mark.parametrize("name", sorted(CHAIN_CASES))
def test_chain_restatement_matches_golden(oracle, name):
    c = CHAIN_CASES[name]
    irs, bands, x = chain_inputs(c)
    y = oracle.chain_run(irs, signals.to_eqband(bands), x, c["sr"], c["block"], FilterSpec(**c["spec"]), makeup=c["makeup"])
    assert np.abs(y - GOLD["chain/" + name]).max() <= TOL


def test_eq_sine_gain_matches_biquad_magnitude(oracle):
    """Invariant the reference's own tests pin (EQProcessorMaxGainTests.cpp:60-86, svfToDisplayBiquad): with sat = 0
    the steady-state sine gain of one SVF band equals the magnitude of the equivalent biquad."""
    sr, f0 = 48000.0, 1000.0
    from oracle.bindings import EqBand
    bands = [EqBand(1000.0, 6.0, 2.0, 1 if i == 8 else 0, 1, 0) for i in range(20)]
    T = 48000
    t = np.arange(T) / sr
    x = 0.1 * np.sin(2 * np.pi * f0 * t)
    l, r, _ = oracle.eq_run(bands, x, x.copy(), sr, 512, saturation=0.0)
    gain = np.abs(l[T // 2:]).max() / 0.1
    assert abs(gain - 10 ** (6.0 / 20.0)) < 2e-3   # peaking band: +6 dB at its centre frequency


def test_epilogue_headroom_and_dither_determinism(oracle):
    x = signals.noise(4096, 1, 0.3)
    y, _, _ = oracle.epilogue(x, 1.25, 48000.0, 0)
    assert np.array_equal(y, (x * 1.25) * 0.8912509381337456)
    u = np.random.default_rng(3).random(2 * 4096)
    q1, tmp, z = oracle.epilogue(x, 1.0, 48000.0, 16, u)
    q2, _, _ = oracle.epilogue(x, 1.0, 48000.0, 16, u)
    assert np.array_equal(q1, q2)
    lsb = 1.0 / 2 ** 15
    assert np.allclose(q1 / lsb, np.round(q1 / lsb))          # quantised to the 16-bit grid
    assert np.abs(q1 - tmp).max() <= 0.5 * lsb + 1e-15         # round-to-nearest of the pre-quantiser value


@pytest.mark.parametrize("name", sorted(DITHER_CASES))
def test_dither_restatement_matches_golden_bit_for_bit(oracle, name):
    """Vectors from PsychoacousticDither.h compiled in place (injected uniforms); chaotic recurrence, so exact equality."""
    c = DITHER_CASES[name]
    x, u = dither_inputs(c)
    q, z = oracle.dither_run(x, u, c["sr"], c["bits"], c["block"])
    assert np.array_equal(q, GOLD["dither/" + name])
    assert np.array_equal(z, GOLD["dither_z/" + name])


@pytest.mark.parametrize("name", sorted(OUTPUT_CASES))
def test_output_stage_restatement_matches_golden(oracle, name):
    """OutputFilter -> makeup -> DC blocker -> headroom -> scrub + clamp against the reference's own OutputFilter.cpp /
    UltraHighRateDCBlocker.h (stereo path: same FMA association, so the restatement is bit-identical)."""
    c = OUTPUT_CASES[name]
    kw = c["kw"]
    got = oracle.output_design(c["sr"], kw.get("conv_is_last", False), kw.get("hc", 1), kw.get("lc", 0), kw.get("lp", 1))
    assert np.array_equal(got, GOLD["output_design/" + name])
    y = oracle.output_run(output_inputs(c), c["sr"], c["block"], **kw)
    assert np.abs(y - GOLD["output/" + name]).max() <= TOL


@pytest.mark.parametrize("name", sorted(FULL_CHAIN_CASES))
def test_full_chain_restatement_matches_golden(oracle, name):
    c = FULL_CHAIN_CASES[name]
    irs, bands, x = chain_inputs(c)
    y = oracle.chain_run(irs, signals.to_eqband(bands), x, c["sr"], c["block"], FilterSpec(**c["spec"]), do_epilogue=False)
    y = oracle.output_run(y, c["sr"], c["block"], makeup=c["makeup"], **c["out"])
    assert np.abs(y - GOLD["full_chain/" + name]).max() <= TOL


def test_output_filter_is_block_size_independent_and_clamps(oracle):
    c = OUTPUT_CASES["eq_last_natural"]
    x = output_inputs(c)
    a = oracle.output_run(x, c["sr"], 512, **c["kw"])
    b = oracle.output_run(x, c["sr"], 64, **c["kw"])
    assert np.array_equal(a, b)
    assert np.abs(a).max() == 0.8912509381337456   # the 3x noise input overshoots: the hard clamp is exercised


def test_outer_mix_restatement_and_peak_latency(oracle):
    """ConvolverProcessor's settled dry/wet mix (restated, unpinned) against its formula, and estimatePeakLatencySamples of the
    oracle against the product's host-only helper (no device needed)."""
    from convopeq_b200.engine import ir_peak_latency
    rng = np.random.default_rng(3)
    wet, dry = rng.standard_normal(4096), rng.standard_normal(4096)
    wet[7] = np.inf
    wet[9] = 2e300
    eps = lambda v: oracle.lib.cpqo_equal_power_sin(float(v))
    for mix, delay in ((1.0, 0), (0.9995, 100), (0.5, 777), (0.0005, 64), (0.0, 512)):
        got = oracle.outer_mix(wet, dry, mix, delay)
        d = np.concatenate([np.zeros(delay), dry[:4096 - delay]])
        m = float(np.float32(mix))
        if m <= 0.001:
            want = d
        else:
            w = np.where(np.isfinite(wet) & (np.abs(wet) < 1e300), wet, 0.0)
            want = w * eps(m) + d * (eps(1.0 - m) if m < 0.999 else 0.0)
        assert np.array_equal(got, want), mix
    assert np.array_equal(oracle.outer_mix(wet, dry, 1.0, 0), oracle.outer_wet(wet, 1.0))
    for n, seed in ((1000, 1), (65536, 2), (30000, 3)):
        a, b = signals.synth_ir(n, seed), np.roll(signals.synth_ir(n, seed + 10), n // 7)
        assert ir_peak_latency(a, b) == oracle.ir_peak_latency(a, b) > 0
        assert ir_peak_latency(a) == oracle.ir_peak_latency(a)
    assert ir_peak_latency(np.zeros(100)) == 0
    one = np.zeros(500)
    one[123] = 1.0
    assert ir_peak_latency(one) == 123


def test_uniform_extension_restatement_is_linear_convolution(oracle):
    """The restatement with the L0 cap lifted (the product's uniform-partition extension, BASELINE config 1 wording) is plain
    linear convolution."""
    from scipy.signal import fftconvolve
    ir, x = signals.synth_ir(40000, 5), signals.noise(16384, 6)
    y, lay = oracle.nuc_run(ir, x, 512, uniform=True)
    assert lay["num_layers"] == 1 and lay["layers"][0]["num_parts_ir"] == 79
    assert np.abs(y - fftconvolve(x, ir)[:16384]).max() <= 1e-13
