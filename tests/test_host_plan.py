"""CPU tests of the host logic behind the C ABI: the library loads and exports what include/cpq.h declares,
the layer plan equals the reference's, and the offline formulation (plain block convolution per layer + the
gather plan) reproduces the callback-by-callback state machine. No compute entry point is called."""
import ctypes as C
import re

import numpy as np
import pytest

from convopeq_b200 import capi
from convopeq_b200.engine import plan_layout, design_band, Band
from oracle.bindings import FilterSpec as OFilterSpec
from tests import signals


def test_library_exports_every_declared_symbol():
    L = capi.load()
    header = open(capi.lib_path().replace("convopeq_b200/libcpq.so", "include/cpq.h")).read()
    declared = set(re.findall(r"\b(cpq_[a-z_0-9]+)\s*\(", header))
    declared -= {"cpq_filter_spec_default"} - set(capi.EXPORTS)
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    for name in capi.EXPORTS:
        assert hasattr(L, name), name
    assert L.cpq_abi_version() == 5


def test_no_cpu_fallback():
    """Without a device every compute path must fail loudly (CPQ_ERR_CUDA), never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = capi.load()
    cfg = capi.Config()
    L.cpq_config_default(C.byref(cfg))
    h = C.c_void_p()
    st = L.cpq_create(C.byref(cfg), C.byref(h))
    assert st == capi.ERR_CUDA and not h.value
    assert b"no CPU fallback" in L.cpq_last_error(None)


def _spec_pair(**kw):
    o = OFilterSpec(**kw)
    c = capi.default_filter_spec(**kw)
    return o, c


PLAN_CASES = [
    (4096, 512, None), (65536, 512, None), (131072, 512, None), (262144, 512, dict(sample_rate=96000.0)),
    (2097152, 512, dict(sample_rate=192000.0)), (65536, 256, {}), (65536, 64, None), (65536, 1024, None),
    (65536, 128, dict(sample_rate=44100.0)), (2047, 256, None), (2048, 256, None), (2049, 256, None),
    (70000, 512, dict(tail_mode=0, tail_start_seconds=0.03)), (70000, 512, dict(tail_mode=2)),
    (70000, 512, dict(tail_enabled=0)), (300000, 512, dict(tail_l1l2_multiplier=16)),
    (300000, 512, dict(tail_l1l2_multiplier=2, tail_strength=1.7)), (100, 512, None), (1, 64, None),
]


@pytest.mark.parametrize("ir_len,block,kw", PLAN_CASES)
def test_layer_plan_matches_checker(checker, ir_len, block, kw):
    ospec = cspec = None
    if kw is not None:
        ospec, cspec = _spec_pair(**kw)
    want = checker.nuc_layout_only(ir_len, block, ospec)
    got, _ = plan_layout(ir_len, block, cspec, 8)
    assert got.num_layers == want["num_layers"]
    for li in range(got.num_layers):
        g, w = got.layers[li], want["layers"][li]
        assert (g.part_size, g.fft_size, g.num_parts_ir, g.num_parts, g.parts_per_callback, g.output_delay_samples) == \
               (w["part_size"], w["fft_size"], w["num_parts_ir"], w["num_parts"], w["parts_per_callback"], w["output_delay_samples"])
        if li > 0:
            assert g.gain == want["gains"][li]


def _offline_numpy(ir, x, block, cspec, scale=1.0):
    """The product's formulation in numpy: per layer a plain linear convolution of x with the layer's IR slice
    (no spectrum filter), then y = y0 + sum g_l * y_l[gather]."""
    n_cb = len(x) // block
    lay, tails = plan_layout(len(ir), block, cspec, n_cb)
    y = np.zeros(len(x))
    for li in range(lay.num_layers):
        l = lay.layers[li]
        h = ir[l.ir_offset:l.ir_offset + l.ir_len] * scale
        full = np.convolve(x, h) if len(h) < 2000 else __import__("scipy.signal").signal.fftconvolve(x, h)
        K = len(x) // l.part_size
        yl = full[:K * l.part_size]
        if li == 0:
            y[:len(yl)] += yl
        else:
            src = tails[li - 1]
            for c in range(n_cb):
                if src[c] >= 0:
                    y[c * block:(c + 1) * block] += l.gain * yl[src[c]:src[c] + block]
    return y, lay


@pytest.mark.parametrize("ir_len,block,T", [(65536, 512, 65536), (65536, 256, 32768), (65536, 64, 16384),
                                            (65536, 1024, 131072), (40000, 128, 32768), (300000, 2048, 262144)])
def test_offline_formulation_equals_state_machine(oracle, ir_len, block, T):
    """Regular and irregular (drop/starve) plans: block convolution + gather plan == Add/Get loop."""
    ir = signals.synth_ir(ir_len, 2)
    x = signals.noise(T, 1)
    want, _ = oracle.nuc_run(ir, x, block)
    got, lay = _offline_numpy(ir, x, block, None)
    assert np.abs(got - want).max() < 1e-12


@pytest.mark.parametrize("known,call,ir_len,n_cb", [(128, 100, 20000, 200), (128, 96, 20000, 200), (512, 441, 140000, 44), (512, 480, 70000, 100),
                                                    (1024, 1000, 140000, 20), (2048, 1125, 70000, 20)])
def test_offline_formulation_for_non_power_of_two_hosts(oracle, known, call, ir_len, n_cb):
    """SetImpulse(known block) with Add/Get calls of the host block: the L0 output ring's per-callback source / count and the tail
    gather reproduce the Add/Get loop -- including the caller's zero-fill of everything beyond what the ring delivered
    (StereoConvolver::process, Runtime.cpp:1174-1183), which blanks the tails too in a short callback."""
    from scipy.signal import fftconvolve
    from convopeq_b200.engine import plan_layout_ex
    T = call * n_cb
    ir, x = signals.synth_ir(ir_len, 3), signals.noise(T, 4)
    want, _ = oracle.nuc_run(ir, x, known, call=call)
    lay, tails, l0s, l0c = plan_layout_ex(ir_len, known, call, None, n_cb)
    y = np.zeros(T)
    for li in range(lay.num_layers):
        l = lay.layers[li]
        full = fftconvolve(x, ir[l.ir_offset:l.ir_offset + l.ir_len])
        K = T // l.part_size
        yl = np.zeros(T + l.part_size)
        yl[:K * l.part_size] = full[:K * l.part_size]
        for c in range(n_cb):
            n = l0c[c]
            if li == 0:
                y[c * call:c * call + n] += yl[l0s[c]:l0s[c] + n]
            elif tails[li - 1][c] >= 0:
                s0 = tails[li - 1][c]
                y[c * call:c * call + n] += l.gain * yl[s0:s0 + n]
    assert np.abs(y - want).max() < 1e-12
    assert (l0c < call).sum() > 1 or call in (96, 480, 1000)     # most of these hosts see recurring short callbacks


def test_irregular_plan_reports_skips():
    lay, tails = plan_layout(65536, 1024, None, 400)
    assert lay.layers[1].skipped_callbacks > 0          # B=1024 default plan drops tail samples
    lay, tails = plan_layout(65536, 512, None, 400)
    assert lay.layers[1].skipped_callbacks == 0 and lay.layers[1].first_output_sample == 7168


def test_design_band_matches_checker(checker):
    for ty in range(5):
        for (f, g, q) in [(1000, 3, 0.7), (19, 50, 25), (30000, -60, 0.001), (100, 0, 1), (25, -6.5, 4)]:
            for sr in (44100.0, 48000.0, 96000.0, 192000.0):
                a = design_band(Band(f, g, q, True, ty, 0), sr)
                b = checker.eq_design(ty, f, g, q, sr)
                got = np.array([a.a1, a.a2, a.a3, a.m0, a.m1, a.m2])
                assert np.allclose(got, b, rtol=4e-16, atol=1e-300), (ty, f, g, q, sr, got, b)


PRESET = """# made-up preset in the EqualizerAPO / AutoEq "ParametricEq.txt" format
Preamp: -5.25 dB   ; trailing comment
Filter 1: ON LSC Fc 105 Hz Gain 4.5 dB Q 0.70
Filter 2: ON PK Fc 63.5 Hz Gain -2.25 dB Q 1.41
Filter 3: OFF PK Fc 250 Hz Gain 3 dB Q 2
Channel: L
Filter 4: ON HSC Fc 9000 Hz Gain -1.5 dB
Filter: ON LP Fc 18000 Hz Q 0.5
channel: R, L
Filter 6: on highpass fc 30 hz q 0   # q <= 0 -> default Q
Channel: Right
Filter 7: ON Fc 1000 Hz Gain 1 dB Q 3.3
Filter 8: ON PK Gain 2 dB Q 1
"""


def test_eq_preset_text_parser():
    """EQProcessor::loadFromTextFile (EQProcessor.Core.cpp:300-495) restated host-side (cpq_parse_eq_preset)."""
    from convopeq_b200.engine import load_eq_preset, EQPARAMETERS_DEFAULT_FREQS as DEFAULT_FREQS
    bands, gain, ignored = load_eq_preset(PRESET)
    f32 = lambda v: float(np.float32(v))
    assert ignored == 0 and gain == f32(-5.25)
    want = [  # (freq, gain, q, enabled, type, mode)
        (105.0, 4.5, 0.70, True, 0, 0), (63.5, -2.25, 1.41, True, 1, 0), (250.0, 3.0, 2.0, False, 1, 0),
        (9000.0, -1.5, 0.707, True, 2, 1), (18000.0, 0.0, 0.5, True, 3, 1), (30.0, 0.0, 0.707, True, 4, 0),
        (1000.0, 1.0, 3.3, True, 1, 2), (DEFAULT_FREQS[7], 2.0, 1.0, True, 1, 2)]
    for b, w in zip(bands, want):
        assert (b.frequency, b.gain, b.q, b.enabled, b.type, b.channel_mode) == (f32(w[0]), f32(w[1]), f32(w[2]), w[3], w[4], w[5]), (b, w)
    for i in range(8, 20):   # untouched bands: disabled, Stereo, 0 dB, frequency kept
        assert not bands[i].enabled and bands[i].gain == 0.0 and bands[i].channel_mode == 0 and bands[i].frequency == DEFAULT_FREQS[i]
    many = "Preamp: -60 dB\n" + "".join(f"Filter {i}: ON PK Fc {100 + i} Hz Gain 1 dB Q 1\n" for i in range(1, 24))
    bands, gain, ignored = load_eq_preset(many)
    assert ignored == 3 and gain == -48.0 and all(b.enabled for b in bands)      # setTotalGain clamps to +-48 dB


def test_eq_preset_reference_fixture():
    """The reference's own sample preset (sampledata/*ParametricEq.txt), when the reference tree is mounted."""
    import glob, os
    from convopeq_b200.engine import load_eq_preset
    files = glob.glob("/root/reference/sampledata/*ParametricEq.txt")
    if not files:
        pytest.skip("reference tree absent")
    text = open(files[0], encoding="utf-8", errors="replace").read()
    bands, gain, ignored = load_eq_preset(text)
    n_filters = sum(1 for ln in text.splitlines() if ln.strip().lower().startswith("filter"))
    assert ignored == max(0, n_filters - 20) and sum(b.enabled for b in bands) == min(n_filters, 20)
    assert gain < 0.0 and bands[0].type == 0 and bands[0].enabled and all(b.q > 0 for b in bands)


def test_ir_scale_factor_matches_restatement(oracle):
    """IRConverter::computeScaleFactor: the product's host helper against the restatement (whose FFT stage is pinned to the
    reference's IRAnalyzer.cpp in tests/test_oracle_vs_ref.py); clamps and jump protection exercised."""
    from convopeq_b200.engine import ir_scale_factor, ir_freq_peak_gain
    t = np.arange(20000)
    ring = np.sin(2 * np.pi * 0.0123 * t) * np.exp(-t / 5000.0)                 # narrow resonance: frequency clamp
    spike = np.zeros(4000); spike[10] = 1.0; spike[500:] = 1e-3                # peak clamp after the energy scale
    cases = [(signals.synth_ir(1000, 1), None), (signals.synth_ir(65536, 2), np.roll(signals.synth_ir(65536, 3), 900)),
             (signals.synth_ir(100000, 4), signals.synth_ir(100000, 5)), (ring, None), (spike, 0.5 * spike), (np.ones(7), None)]
    for a, b in cases:
        g, w = ir_freq_peak_gain(a, b), oracle.ir_freq_peak_gain(a, b)
        assert abs(g - w) <= 1e-10 * w
        got, want = ir_scale_factor(a, b), oracle.ir_scale_factor(a, b)
        assert got[1] and want[1] and abs(got[0] - want[0]) <= 1e-10 * want[0] and abs(got[2] - want[2]) <= 1e-4
    assert ir_scale_factor(ring)[2] > 50.0 and ir_scale_factor(spike)[2] > 0.0
    # jump protection against the playing IR: after the peak / RMS clamps the new IR never exceeds 0.5 / 0.25, which are
    # also the protection's absolute thresholds, so it cannot lower the scale further -- same answer with and without
    loud = signals.synth_ir(3000, 6) * 50.0
    quiet = signals.synth_ir(3000, 7)
    free = ir_scale_factor(loud)
    held = ir_scale_factor(loud, None, quiet, None, 0.01)
    want = oracle.ir_scale_factor(loud, None, quiet, None, 0.01)
    assert abs(held[0] - want[0]) <= 1e-10 * want[0] and abs(held[0] - free[0]) <= 1e-12 * free[0]
    assert ir_scale_factor(np.zeros(100)) == (1.0, True, 0.0)


def test_ir_minimum_phase_conversion():
    """convertToMinimumPhase (ResampleAndFallback.cpp:333-460) restated host-side: same steps in numpy, plus what a
    minimum-phase version must satisfy (same magnitude response, energy moved to the front)."""
    from convopeq_b200.engine import ir_min_phase
    for n, seed in ((300, 1), (4096, 2), (20000, 3)):
        ir = np.roll(signals.synth_ir(n, seed), n // 3)          # delayed: far from minimum phase
        got = ir_min_phase(ir)
        N = 1 << int(np.ceil(np.log2(4 * n)))
        X = np.fft.fft(ir, N)
        c = np.fft.ifft(np.log(np.maximum(np.abs(X), 1e-300))).real
        fold = np.zeros(N)
        fold[0], fold[N // 2] = c[0], c[N // 2]
        fold[1:N // 2] = 2.0 * c[1:N // 2]
        S = np.fft.fft(fold)
        want = np.fft.ifft(np.exp(np.clip(S.real, -50, 50) + 1j * np.clip(S.imag, -50, 50))).real[:n]
        want[np.abs(want) < 1e-18] = 0.0
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
        G = np.abs(np.fft.rfft(got, N))
        assert np.abs(G - np.abs(X[:N // 2 + 1])).max() <= 2e-2 * np.abs(X).max()     # truncation to n samples only
        e_in, e_out = np.cumsum(ir ** 2), np.cumsum(got ** 2)
        assert np.all(e_out[: n // 2] >= e_in[: n // 2] - 1e-12) and e_out[n // 8] > 0.5 * e_out[-1]
    assert ir_min_phase(np.zeros(3 * 1024 * 1024)) is None      # FFT would exceed 2^23 points: the reference skips too


def test_ir_prepare_matches_restatement_and_has_the_loader_shape(oracle):
    """cpq_ir_prepare (host-only) = LoaderThread::doLoadStep after the resampler: DC blocker, asymmetric Tukey, trim + fade."""
    from convopeq_b200.engine import ir_prepare
    from tests import signals
    for sr, n, secs in ((48000.0, 30000, 1.0), (96000.0, 200000, 1.5), (44100.0, 5000, 0.05), (48000.0, 700, 3.0)):
        ir = signals.synth_ir(n, 3)
        ir[n // 50] = 1.5                      # a clear peak away from the start: both tapers are exercised
        got = ir_prepare(ir, sr, secs)
        want = oracle.ir_prepare(ir, sr, secs)
        target = min(max(int(sr * float(np.float32(secs))), 1), 2097152)
        assert got.size == target == want.size
        assert np.abs(got - want).max() <= 1e-15
        copy = min(target, n)
        assert np.all(got[copy:] == 0.0)
        if copy > 300:
            assert abs(got[copy - 1]) <= abs(ir).max() * 2.0 / 256   # the fade-out reaches (almost) zero at the last copied sample
        if (n // 50) * 0.05 >= 1.0:
            assert abs(got[0]) <= 1e-15            # the pre-taper starts at zero gain
