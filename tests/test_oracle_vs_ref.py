"""The C restatement against the reference's own translation units (oracle/_ref), where they were compiled.
Skipped where the reference tree is absent (the golden vectors pin the restatement there)."""
import numpy as np
import pytest

from oracle.bindings import FilterSpec
from tests import signals

CASES = [
    (4096, 512, None, 16384), (65536, 512, None, 32768), (65536, 512, {}, 32768), (65536, 256, {}, 16384),
    (65536, 64, None, 8192), (65536, 1024, None, 65536), (30000, 480, None, 24000), (2049, 256, dict(sample_rate=96000.0), 4096),
    (70000, 512, dict(tail_mode=0, tail_start_seconds=0.03), 32768), (70000, 512, dict(tail_enabled=0), 16384),
    (300000, 512, dict(tail_l1l2_multiplier=2, tail_strength=1.7), 65536), (50, 128, None, 1024),
    (65536, 512, dict(hc_mode=0, lc_mode=1), 16384), (65536, 512, dict(hc_mode=2, sample_rate=44100.0), 16384),
]


@pytest.mark.parametrize("ir_len,block,kw,T", CASES)
def test_nuc_restatement_equals_reference(oracle, ref, ir_len, block, kw, T):
    ir = signals.synth_ir(ir_len, 2)
    x = signals.noise(T, 1)
    spec = FilterSpec(**kw) if kw is not None else None
    yo, lo = oracle.nuc_run(ir, x, block, spec=spec)
    yr, lr = ref.nuc_run(ir, x, block, spec=spec)
    assert lo == lr
    assert np.abs(yo - yr).max() <= 1e-13


def test_nuc_spectra_equal_reference(oracle, ref):
    ir = signals.synth_ir(70000, 4)
    for kw in (None, {}, dict(tail_mode=0)):
        spec = FilterSpec(**kw) if kw is not None else None
        so, _ = oracle.nuc_spectra(ir, 512, 0.5, spec)
        sr_, _ = ref.nuc_spectra(ir, 512, 0.5, spec)
        for a, b in zip(so, sr_):
            assert np.abs(a - b).max() <= 1e-13


@pytest.mark.parametrize("ir_len,block,spec,call", [(65536, 512, None, None), (20000, 512, {}, None), (20, 64, {}, None),
                                                   (9000, 256, None, 100), (65536, 1024, {}, None), (40000, 128, dict(tail_mode=0), None)])
def test_direct_head_restatement_equals_reference(oracle, ref, ir_len, block, spec, call):
    """enableDirectHead = true: direct FIR of the first <= 32 taps + partitions built without them."""
    ir, x = signals.synth_ir(ir_len, 3), signals.noise(8192, 4)
    sp = FilterSpec(**spec) if spec is not None else None
    yo, _ = oracle.nuc_run(ir, x, block, 0.7, sp, call, direct_head=True)
    yr, _ = ref.nuc_run(ir, x, block, 0.7, sp, call, direct_head=True)
    assert np.abs(yo - yr).max() <= 1e-13
    if spec is not None:   # the head bypasses the spectrum filter: a different result, not a rounding difference
        yp, _ = ref.nuc_run(ir, x, block, 0.7, sp, call, direct_head=False)
        assert np.abs(yr - yp).max() > 1e-6


def test_non_block_sized_calls(oracle, ref):
    """Add/Get with call sizes different from the prepared block size (ring latency path)."""
    ir = signals.synth_ir(20000, 5)
    x = signals.noise(9600, 3)
    for call in (480, 100, 512):
        yo, _ = oracle.nuc_run(ir, x, 512, call=call)
        yr, _ = ref.nuc_run(ir, x, 512, call=call)
        assert np.abs(yo - yr).max() <= 1e-13


@pytest.mark.parametrize("name,bkw,kw", [
    ("default", dict(seed=7), {}), ("sat0", dict(seed=7), dict(saturation=0.0)), ("stress", dict(seed=7, stress=True), {}),
    ("modes", dict(seed=8, modes=[i % 3 for i in range(20)]), {}), ("types", dict(seed=9, types=[i % 5 for i in range(20)]), {}),
    ("gain", dict(seed=7), dict(total_gain_db=-3.0)), ("ramp", dict(seed=7), dict(gain_change_db=-6.0, gain_change_at=512 * 20)),
])
def test_eq_restatement_equals_reference(oracle, ref, name, bkw, kw):
    sr, T = 48000.0, 48000
    T = T // 512 * 512
    bands = signals.to_eqband(signals.band_params(**bkw))
    xl, xr = signals.log_sweep(T, sr)
    lo, ro, so = oracle.eq_run(bands, xl, xr, sr, 512, **kw)
    lr, rr, s_r = ref.eq_run(bands, xl, xr, sr, 512, **kw)
    scale = max(1.0, np.abs(lr).max())
    assert np.abs(lo - lr).max() <= 1e-13 * scale and np.abs(ro - rr).max() <= 1e-13 * scale
    assert np.abs(so - s_r).max() <= 1e-12 * scale


@pytest.mark.parametrize("name,bkw,kw", [
    ("parallel", dict(seed=7), dict(structure=1)),
    ("parallel_lr_sat0", dict(seed=8, modes=[i % 3 for i in range(20)]), dict(structure=1, saturation=0.0)),
    ("agc", dict(seed=7), dict(agc=True)),
    ("agc_parallel", dict(seed=11), dict(agc=True, structure=1)),
    ("mid_side", dict(seed=7, modes=[0, 3, 4, 1, 2] * 4, flat=[5, 6, 7]), {}),
    ("mid_side_agc", dict(seed=10, modes=[3, 4] * 10), dict(agc=True)),
    ("parallel_mid_side", dict(seed=7, modes=[0, 3, 4, 1, 2] * 4, flat=[5, 6]), dict(structure=1)),
    ("parallel_mid_side_agc", dict(seed=9, modes=[3, 4] * 10), dict(structure=1, agc=True)),
])
@pytest.mark.parametrize("block", [64, 512, 1000])
def test_eq_modes_restatement_equals_reference(oracle, ref, name, bkw, kw, block):
    """Parallel structure, AGC and Mid/Side bands of the restatement against the reference's own
    process(block, params, cache) / process(block)."""
    sr, T = 48000.0, 24000
    bands = signals.to_eqband(signals.band_params(**bkw))
    xl, xr = signals.log_sweep(T, sr)
    xl = 1.7 * xl + signals.noise(T, 5, 0.2)
    xr = 0.4 * xr
    lo, ro, _ = oracle.eq_run(bands, xl, xr, sr, block, **kw)
    lr, rr, _ = ref.eq_run(bands, xl, xr, sr, block, **kw)
    scale = max(1.0, np.abs(lr).max())
    assert np.abs(lo - lr).max() <= 1e-13 * scale and np.abs(ro - rr).max() <= 1e-13 * scale
    assert np.abs(lr - xl).max() > 1e-3      # the EQ did something


def test_node_path_skips_flat_bands(ref, oracle):
    """A Mid/Side band switches the reference to BandNode::active, which drops shelf/peaking bands within 0.01 dB of flat
    (EQProcessor.Coefficients.cpp:49-53) -- unlike EQCoeffCache::bandActive.  The outputs differ measurably."""
    sr, T = 48000.0, 8192
    xl, xr = signals.log_sweep(T, sr)
    p_ms = signals.band_params(seed=7, modes=[3] + [0] * 19, flat=[5])
    p_st = signals.band_params(seed=7, flat=[5])
    for p in (p_ms, p_st):
        a, b = oracle.eq_run(signals.to_eqband(p), xl, xr, sr, 512), ref.eq_run(signals.to_eqband(p), xl, xr, sr, 512)
        assert np.abs(a[0] - b[0]).max() <= 1e-13 and np.abs(a[1] - b[1]).max() <= 1e-13


def test_eq_is_block_size_independent(ref):
    """with a settled gain ramp the reference EQ output does not depend on the callback size."""
    sr, T = 48000.0, 9216
    bands = signals.to_eqband(signals.band_params(7))
    xl, xr = signals.log_sweep(T, sr)
    outs = [ref.eq_run(bands, xl, xr, sr, b)[0] for b in (64, 512, 4608)]
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


@pytest.mark.parametrize("kw", [dict(), dict(conv_is_last=True, hc=0, lc=1), dict(conv_is_last=True, hc=2), dict(lp=0, makeup=1.7),
                                dict(use_filter=False), dict(dc_cutoff=0.0, clamp=False, headroom=False)])
@pytest.mark.parametrize("sr", [44100.0, 48000.0, 192000.0])
def test_output_stage_restatement_equals_reference(oracle, ref, kw, sr):
    """OutputFilter.cpp + UltraHighRateDCBlocker.h compiled from the reference tree vs oracle/cpq_oracle.c."""
    x = np.stack([signals.noise(24000, 31), signals.noise(24000, 32)]) * 2.0 + 0.03
    for m in range(3):
        assert np.array_equal(oracle.output_design(sr, True, m, m % 2, m), ref.output_design(sr, True, m, m % 2, m))
        assert np.array_equal(oracle.output_design(sr, False, m, m % 2, m), ref.output_design(sr, False, m, m % 2, m))
    a, b = oracle.output_run(x, sr, 480, **kw), ref.output_run(x, sr, 480, **kw)
    assert np.abs(a - b).max() <= 1e-13
    # mono path of the reference uses the scalar BiquadState::process (different association): 1e-16-level differences
    # amplified by the 20 Hz high-pass pole radius
    am, bm = oracle.output_run(x[:1], sr, 480, **kw), ref.output_run(x[:1], sr, 480, **kw)
    assert np.abs(am - bm).max() <= 1e-11


@pytest.mark.parametrize("kw", [dict(limiter_ms=100.0), dict(limiter_ms=100.0, use_filter=False, dc_cutoff=0.0), dict(limiter_ms=30.0, clamp=False)])
def test_peak_limiter_restatement_equals_reference(oracle, ref, kw):
    """audioengine/SimplePeakLimiter.h compiled from the reference tree, in processOutputDouble's order (scrub -> limiter -> clamp)."""
    x = np.stack([signals.noise(24000, 31), signals.noise(24000, 32)]) * 6.0
    x[:, 8000:12000] *= 0.05
    a, b = oracle.output_run(x, 48000.0, 480, **kw), ref.output_run(x, 48000.0, 480, **kw)
    assert np.abs(a - b).max() <= 1e-13
    assert np.abs(b - ref.output_run(x, 48000.0, 480, **{**kw, "limiter_ms": 0.0})).max() > 0.1


def test_ir_frequency_peak_gain_restatement_equals_reference(oracle, ref):
    """IRAnalyzer::estimateMaxFrequencyResponseGain (src/IRAnalyzer.cpp compiled in place) -- the FFT stage of
    IRConverter::computeScaleFactor."""
    t = np.arange(20000)
    ring = np.sin(2 * np.pi * 0.0123 * t) * np.exp(-t / 5000.0)
    for a, b in [(signals.synth_ir(1000, 1), None), (signals.synth_ir(65536, 2), np.roll(signals.synth_ir(65536, 3), 900)),
                 (signals.synth_ir(100000, 4), signals.synth_ir(100000, 5)), (signals.synth_ir(3, 6), None), (ring, None)]:
        x, y = oracle.ir_freq_peak_gain(a, b), ref.ir_freq_peak_gain(a, b)
        assert abs(x - y) <= 1e-12 * y


def test_input_transform_restatement_equals_reference(oracle, ref):
    """convo::input_transform::convertDoubleToDoubleHighQuality (src/InputBitDepthTransform.h compiled in place)."""
    x = np.random.default_rng(1).standard_normal(1027) * 0.8
    x[3], x[10], x[11], x[20], x[21], x[30], x[1025], x[1026] = np.nan, np.inf, -np.inf, 1e-25, -3e-21, 2.5, np.inf, np.nan
    for g in (1.0, 0.5, 1.0 + 1e-10, 3.0):
        assert np.array_equal(oracle.input_transform(x, g), ref.input_transform(x, g))


@pytest.mark.parametrize("sr,bits,T,block", [(48000.0, 24, 200000, 512), (96000.0, 16, 70000, 480), (44100.0, 32, 40000, 64),
                                             (192000.0, 20, 30000, 1000)])
def test_dither_restatement_is_bit_identical_to_reference(oracle, ref, sr, bits, T, block):
    """PsychoacousticDither::processStereoBlock compiled in place (mkl_vsl.h shim handing out injected uniforms) against
    cpqo_epilogue_ex: the error-feedback recurrence is chaotic, so only bit identity is a meaningful pin.  200 000 samples
    cross the 65 536-entry ring three times (prefill + refillRandomRingNonRt)."""
    x = np.stack([signals.noise(T, 1, 0.3), signals.noise(T, 2, 0.3)])
    u = np.random.default_rng(5).random((2, 2 * T))
    q, z = ref.dither_run(x, u, sr, bits, block)
    want, zo = oracle.dither_run(x, u, sr, bits, block)
    assert np.array_equal(q, want) and np.array_equal(z, zo)
    qm, zm = ref.dither_run(x[:1], u[:1], sr, bits, block)       # mono path (:362-405)
    wm, zom = oracle.dither_run(x[:1], u[:1], sr, bits, block)
    assert np.array_equal(qm, wm) and np.array_equal(zm, zom)
    lsb = 1.0 / 2 ** (bits - 1)
    assert np.allclose(q / lsb, np.round(q / lsb))


def test_ir_dc_blocker_stage_is_bit_identical_to_reference(oracle, ref):
    """UltraHighRateDCBlocker.h compiled in place at the loader's 1 Hz (LoaderThread.cpp:590-598) against the restated stage."""
    import ctypes as C
    for sr, n in ((48000.0, 50000), (192000.0, 20000)):
        x = signals.synth_ir(n, 5) + 0.01
        want = ref.ir_dc_block(x, sr, 1.0)
        d = np.ascontiguousarray(x).copy()
        f = oracle.lib.cpqo_ir_dc_block
        f.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double, C.c_double]
        f.restype = None
        f(d.ctypes.data_as(C.POINTER(C.c_double)), n, sr, 1.0)
        assert np.array_equal(d, want)


@pytest.mark.parametrize("seed", [1, 0xDEADBEEFCAFEF00D])
def test_dither_fallback_generator_is_bit_identical_to_reference(oracle, ref, seed):
    """PsychoacousticDither(seed) with vslNewStream failing: every uniform comes from the header's own xorshift64* generator
    (fallbackUniform :485-497, seeded through SplitMix64 :118-140).  The restated generator + shaper equal it bit for bit."""
    T = 60000
    for nch in (2, 1):
        x = np.stack([signals.noise(T, 1 + i, 0.3) for i in range(nch)])
        q, z = ref.dither_run_seeded(x, seed, 48000.0, 24, 512)
        w, zo = oracle.dither_run_seeded(x, seed, 48000.0, 24, 512)
        assert np.array_equal(q, w) and np.array_equal(z, zo)
