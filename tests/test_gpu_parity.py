"""GPU parity: the CUDA path through the C ABI against the checker (the compiled reference when
oracle/_ref exists, else the C restatement) on the same seeded inputs.  Tolerance: max abs error
<= 1e-10 of full scale (BASELINE.json north_star); onset positions sample-exact."""
import ctypes as C

import numpy as np
import pytest

from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
from oracle.bindings import FilterSpec as OFilterSpec
from tests import signals

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _specs(kw):
    if kw is None:
        return None, None
    return OFilterSpec(**kw), capi.default_filter_spec(**kw)


def _run_conv(ir_l, ir_r, x, block, sr, kw, boundary=capi.CONV_INNER, scale=1.0):
    ospec, cspec = _specs(kw)
    T = x.shape[1]
    eng = ConvoPeqEngine(n_streams=1, n_channels=2, sample_rate=sr, block_size=block, max_samples=T, conv_boundary=boundary)
    eng.set_impulse(0, 0, ir_l, scale, cspec)
    eng.set_impulse(0, 1, ir_r, scale, cspec)
    y = x.copy()
    eng.process(y, capi.STAGE_CONV)
    lay = eng.layout()
    eng.close()
    return y, lay, ospec


CONV_CASES = [
    # (ir_len, block, T, sr, spec kwargs)            config
    (4096, 512, 16384, 48000.0, None),               # single layer == linear convolution
    (65536, 512, 65536, 48000.0, None),              # cfg1a: L0 12x512 + L1 15x4096, D1 7168, g1 1.4375
    (65536, 512, 65536, 48000.0, {}),                # cfg1a with the production default FilterSpec
    (131072, 512, 49152, 48000.0, {}),               # cfg4 geometry
    (65536, 256, 32768, 48000.0, {}),                # three layers 256/2048/16384 -> needs P=16384: skipped below
    (65536, 64, 16384, 48000.0, None),               # three layers 64/512/4096
    (65536, 1024, 131072, 48000.0, None),            # irregular plan: drops + starves
    (2047, 256, 4096, 48000.0, None), (2048, 256, 4096, 48000.0, None), (2049, 256, 4096, 48000.0, None),
    (70000, 512, 65536, 48000.0, dict(tail_mode=0, tail_start_seconds=0.03)),   # air absorption tilt
    (70000, 512, 32768, 48000.0, dict(tail_mode=2)),
    (262144, 512, 98304, 96000.0, dict(sample_rate=96000.0)),                   # cfg3 geometry
    (262144, 256, 65536, 96000.0, dict(sample_rate=96000.0)),                   # cfg3 at B=256: 256 / 2048 / 16384 (large-FFT path)
    (300000, 2048, 262144, 48000.0, None),                                      # P1 = 16384, irregular plan
    (2097152, 512, 131072, 192000.0, dict(sample_rate=192000.0)),               # cfg5 geometry: L2 56 x 32768, D2 60416, g2 1.1
    (1200000, 1024, 262144, 48000.0, None),                                     # P2 = 65536 (FFT 131072)
]


@pytest.mark.parametrize("ir_len,block,T,sr,kw", CONV_CASES)
def test_convolver_matches_reference(checker, ir_len, block, T, sr, kw):
    ir_l, ir_r = signals.synth_ir(ir_len, 2), signals.synth_ir(ir_len, 3)
    x = np.stack([signals.noise(T, 1), signals.noise(T, 11)])
    y, _, ospec = _run_conv(ir_l, ir_r, x, block, sr, kw)
    for ch, ir in enumerate((ir_l, ir_r)):
        want, _ = checker.nuc_run(ir, x[ch], block, spec=ospec)
        err = np.abs(y[ch] - want).max()
        assert err <= TOL, (ch, err)


def test_impulse_onsets_are_sample_exact(checker):
    """delta at n=0 and n=B-1: the layout (L0 at 0 latency, tails at D_l with gain g_l) is sample exact."""
    ir_len, block, T = 65536, 512, 90112
    ir = signals.synth_ir(ir_len, 5)
    for at in (0, block - 1):
        x = np.stack([signals.impulse(T, at), signals.silence_then_step(T, 1000)])
        y, lay, _ = _run_conv(ir, ir, x, block, 48000.0, None)
        want0, _ = checker.nuc_run(ir, x[0], block)
        want1, _ = checker.nuc_run(ir, x[1], block)
        assert np.abs(y[0] - want0).max() <= TOL and np.abs(y[1] - want1).max() <= TOL
        nz_got = np.flatnonzero(np.abs(y[0]) > 1e-13)
        nz_want = np.flatnonzero(np.abs(want0) > 1e-13)
        assert nz_got[0] == nz_want[0] == at
        assert lay.layers[1].first_output_sample == 7168


def test_outer_boundary_wet_gain(checker, oracle):
    ir = signals.synth_ir(8192, 4)
    T = 16384
    x = np.stack([signals.noise(T, 1), signals.noise(T, 2)])
    y, _, _ = _run_conv(ir, ir, x, 512, 48000.0, None, boundary=capi.CONV_OUTER)
    want, _ = checker.nuc_run(ir, x[0], 512)
    want = oracle.outer_wet(want, 1.0)
    assert np.abs(y[0] - want).max() <= TOL


@pytest.mark.parametrize("mix", [1.0, 0.9995, 0.35, 0.0005, 0.0])
def test_outer_dry_wet_mix(checker, oracle, mix):
    """ConvolverProcessor::process with mix < 1 (settled smoothers): wet * sin-gain(mix) + latency-compensated
    dry * sin-gain(1 - mix); dry-only fast path at mix <= 0.001.  The outer boundary is restated only (parity unpinned)."""
    from convopeq_b200.engine import ir_peak_latency
    sr, block, T, ir_len = 48000.0, 512, 32768, 30000
    irs = [signals.synth_ir(ir_len, 70), np.roll(signals.synth_ir(ir_len, 71), 333)]
    x = np.stack([signals.noise(T, 72), signals.noise(T, 73)])
    eng = ConvoPeqEngine(1, 2, sr, block, T, conv_boundary=capi.CONV_OUTER)
    for ch in range(2):
        eng.set_impulse(0, ch, irs[ch], 1.0, None)
    delay = eng.latency() + ir_peak_latency(irs[0], irs[1])
    assert delay == block + oracle.ir_peak_latency(irs[0], irs[1]) and delay > block
    eng.set_mix(mix, delay)
    eng.set_eq(0, signals.to_band(signals.band_params(74)))
    y = x.copy()
    eng.process(y, capi.STAGE_CONV)
    z = x.copy()
    eng.process(z, capi.STAGE_CONV | capi.STAGE_EQ)
    eng.close()
    w = []
    for ch in range(2):
        wet, _ = checker.nuc_run(irs[ch], x[ch], block)
        w.append(oracle.outer_mix(wet, x[ch], mix, delay))
        assert np.abs(y[ch] - w[ch]).max() <= TOL, (ch, mix)
    if mix <= 0.001:
        assert np.array_equal(y[0][delay:], x[0][:T - delay]) and not y[0][:delay].any()
    if mix == 0.35:
        # bypassed convolver (processBypassWithLatencyCompensation, Runtime.cpp:123-186): the delayed input whatever the mix,
        # followed by the EQ as usual
        eng2 = ConvoPeqEngine(1, 2, sr, block, T, conv_boundary=capi.CONV_OUTER)
        for ch in range(2):
            eng2.set_impulse(0, ch, irs[ch], 1.0, None)
        eng2.set_mix(mix, delay)
        eng2.set_convolver_bypass(True)
        eng2.set_eq(0, signals.to_band(signals.band_params(74)))
        b = x.copy()
        eng2.process(b, capi.STAGE_CONV)
        assert np.array_equal(b[:, delay:], x[:, :T - delay]) and not b[:, :delay].any()
        b2 = x.copy()
        eng2.process(b2, capi.STAGE_CONV | capi.STAGE_EQ)
        eng2.close()
        bl, br, _ = checker.eq_run(signals.to_eqband(signals.band_params(74)), b[0], b[1], sr, block)
        assert np.abs(b2[0] - bl).max() <= TOL and np.abs(b2[1] - br).max() <= TOL
    wl, wr, _ = checker.eq_run(signals.to_eqband(signals.band_params(74)), w[0], w[1], sr, block)
    assert np.abs(z[0] - wl).max() <= TOL and np.abs(z[1] - wr).max() <= TOL


@pytest.mark.parametrize("ir_len,block,T,kw,shared", [(65536, 512, 32768, {}, False), (65536, 512, 16384, None, True), (20, 64, 4096, {}, False),
                                                      (131072, 1024, 65536, dict(tail_mode=0), False)])
def test_direct_head_matches_reference(checker, ir_len, block, T, kw, shared):
    """SetImpulse(..., enableDirectHead = true): the first <= 32 taps as a direct-form FIR outside the spectrum filter."""
    ospec, cspec = _specs(kw)
    n_streams = 3
    irs = [signals.synth_ir(ir_len, 80 + i) for i in range(2 if shared else 2 * n_streams)]
    x = np.stack([signals.noise(T, 90 + i) for i in range(2 * n_streams)])
    eng = ConvoPeqEngine(n_streams, 2, 48000.0, block, T, shared_ir=shared)
    eng.set_direct_head(True)
    for s in range(1 if shared else n_streams):
        for ch in range(2):
            eng.set_impulse(-1 if shared else s, ch, irs[2 * s + ch], 0.7, cspec)
    with pytest.raises(capi.CpqError):
        eng.set_direct_head(False)          # the head was already removed from the partitions
    y = x.copy()
    eng.process(y, capi.STAGE_CONV)
    eng.close()
    for q in range(2 * n_streams):
        ir = irs[q % 2] if shared else irs[q]
        want, _ = checker.nuc_run(ir, x[q], block, scale=0.7, spec=ospec, direct_head=True)
        assert np.abs(y[q] - want).max() <= TOL, q


def test_uniform_partition_extension(oracle):
    """BASELINE config 1 wording, "uniform partitioned convolution block 512" for 65,536 taps = 128 x 512: not expressible in the
    reference (L0 is capped at 32 partitions), offered as cfg.uniform_partitions.  Checked against linear convolution and
    against the restatement with the same cap lifted."""
    from scipy.signal import fftconvolve
    ir_len, block, T = 65536, 512, 65536
    irs = [signals.synth_ir(ir_len, 2), signals.synth_ir(ir_len, 3)]
    x = np.stack([signals.noise(T, 1), signals.noise(T, 11)])
    eng = ConvoPeqEngine(1, 2, 48000.0, block, T, uniform_partitions=True)
    for ch in range(2):
        eng.set_impulse(0, ch, irs[ch], 1.0, None)
    lay = eng.layout()
    assert lay.num_layers == 1 and lay.layers[0].num_parts_ir == 128 and lay.layers[0].part_size == 512
    y = x.copy()
    eng.process(y, capi.STAGE_CONV)
    eng.close()
    for ch in range(2):
        lin = fftconvolve(x[ch], irs[ch])[:T]
        assert np.abs(y[ch] - lin).max() <= 1e-12 * max(1.0, np.abs(lin).max())
        want, lo = oracle.nuc_run(irs[ch], x[ch], block, uniform=True)
        assert lo["num_layers"] == 1 and np.abs(y[ch] - want).max() <= TOL


@pytest.mark.parametrize("block,ir_len,T,kw", [(480, 65536, 480 * 80, None), (480, 131072, 480 * 120, {}), (96, 30000, 96 * 300, None),
                                               (441, 65536, 441 * 100, {}), (1000, 131072, 1000 * 50, None),
                                               (960, 70000, 960 * 60, {}), (1920, 300000, 1920 * 40, None), (480, 300, 480 * 10, None)])
def test_non_power_of_two_host_block(checker, block, ir_len, T, kw):
    """Hosts whose block is not a power of two (480 = 10 ms at 48 kHz, ...): the application prepares the convolver with the block
    rounded up (SetImpulse(knownBlockSize)) and calls Add/Get with the host block, so L0 goes through its output ring with
    the reference's varying latency and the tails follow the per-call distribution schedule."""
    ospec, cspec = _specs(kw)
    known = 64
    while known < block:
        known *= 2
    irs = [signals.synth_ir(ir_len, 60 + ch) for ch in range(2)]
    x = np.stack([signals.noise(T, 62 + ch) for ch in range(2)])
    eng = ConvoPeqEngine(1, 2, 48000.0, block, T)
    for ch in range(2):
        eng.set_impulse(0, ch, irs[ch], 1.0, cspec)
    assert eng.latency() == known
    y = x.copy()
    eng.process(y, capi.STAGE_CONV)
    eng.close()
    for ch in range(2):
        want, _ = checker.nuc_run(irs[ch], x[ch], known, spec=ospec, call=block)
        assert np.abs(y[ch] - want).max() <= TOL, ch
        assert np.abs(want).max() > 1e-3


@pytest.mark.parametrize("n_callbacks,f32", [(1, False), (101, False), (37, True)])
def test_odd_number_of_samples_per_call(checker, n_callbacks, f32):
    """A 441-sample host may hand over an odd number of samples (one callback, 101 callbacks): rows keep their 16-byte alignment
    through an even pitch and the paired accesses take the pad sample along.  Conv -> EQ (AGC) -> epilogue; FP64 and FP32 host
    buffers."""
    block, T, sr = 441, 441 * n_callbacks, 48000.0
    irs = [signals.synth_ir(9000, 910 + ch) for ch in range(2)]
    x = np.stack([signals.noise(T, 912 + ch, 0.3) for ch in range(2)])
    p = signals.band_params(913)
    eng = ConvoPeqEngine(1, 2, sr, block, T, conv_boundary=capi.CONV_OUTER)
    for ch in range(2):
        eng.set_impulse(0, ch, irs[ch], 1.0, capi.default_filter_spec())
    eng.set_eq(0, signals.to_band(p), agc=True)
    eng.set_epilogue(1.1, 0)
    if f32:
        y32 = x.astype(np.float32)
        xin = y32.astype(np.float64)
        eng.process_f32(y32, capi.STAGE_ALL)
        y, tol = y32.astype(np.float64), 1e-6
    else:
        xin, y, tol = x, x.copy(), TOL
        eng.process(y, capi.STAGE_ALL)
    eng.close()
    want = checker.chain_run((irs[0], irs[1]), signals.to_eqband(p), xin, sr, block, OFilterSpec(), makeup=1.1, agc=True, known_block=512)
    assert np.isfinite(y).all() and np.abs(y - want).max() <= tol


@pytest.mark.parametrize("block", [480, 441])
@pytest.mark.parametrize("agc", [False, True])
def test_non_power_of_two_host_block_full_chain(checker, agc, block):
    """480- / 441-sample host callbacks through conv -> EQ (AGC per callback: a thread's 32 samples may straddle two callbacks)
    -> epilogue, three streams."""
    sr, T, ir_len, n = 48000.0, block * 100, 131072, 3
    eng = ConvoPeqEngine(n, 2, sr, block, T, conv_boundary=capi.CONV_OUTER)
    x = np.stack([signals.noise(T, 400 + i, 0.3) for i in range(2 * n)])
    irs = [signals.synth_ir(ir_len, 420 + i) for i in range(2 * n)]
    for s in range(n):
        for ch in range(2):
            eng.set_impulse(s, ch, irs[2 * s + ch], 1.0, capi.default_filter_spec())
        eng.set_eq(s, signals.to_band(signals.band_params(440 + s)), agc=agc)
    eng.set_epilogue(1.1, 0)
    y = x.copy()
    eng.process(y, capi.STAGE_ALL)
    eng.close()
    for s in range(n):
        want = checker.chain_run((irs[2 * s], irs[2 * s + 1]), signals.to_eqband(signals.band_params(440 + s)), x[2 * s:2 * s + 2], sr, block,
                                 OFilterSpec(), makeup=1.1, agc=agc, known_block=512)
        assert np.abs(y[2 * s:2 * s + 2] - want).max() <= TOL, s


def test_ir_scale(checker):
    ir = signals.synth_ir(20000, 4)
    T = 16384
    x = np.stack([signals.noise(T, 1), signals.noise(T, 2)])
    y, _, _ = _run_conv(ir, ir, x, 512, 48000.0, {}, scale=0.37)
    want, _ = checker.nuc_run(ir, x[1], 512, scale=0.37, spec=OFilterSpec())
    assert np.abs(y[1] - want).max() <= TOL


EQ_CASES = [
    ("default", dict(seed=7), dict()),
    ("sat0", dict(seed=7), dict(saturation=0.0)),
    ("stress_q20", dict(seed=7, stress=True), dict()),
    ("modes", dict(seed=8, modes=[i % 3 for i in range(20)]), dict()),
    ("types", dict(seed=9, types=[i % 5 for i in range(20)]), dict()),
    ("disabled", dict(seed=10, enabled=[i % 2 for i in range(20)]), dict()),
    ("gain", dict(seed=7), dict(total_gain_db=-3.0)),
]


@pytest.mark.parametrize("name,bkw,kw", EQ_CASES)
@pytest.mark.parametrize("T", [4096 * 3 + 512, 96000])
def test_eq_matches_reference(checker, name, bkw, kw, T):
    sr, block = 48000.0, 512
    T = T // block * block
    params = signals.band_params(**bkw)
    xl, xr = signals.log_sweep(T, sr)
    eng = ConvoPeqEngine(1, 2, sr, block, T)
    eng.set_eq(0, signals.to_band(params), kw.get("saturation", 0.2), kw.get("total_gain_db", 0.0))
    y = np.stack([xl, xr]).copy()
    eng.process(y, capi.STAGE_EQ)
    state = eng.eq_state(0)
    eng.close()
    wl, wr, wstate = checker.eq_run(signals.to_eqband(params), xl, xr, sr, block, **kw)
    assert np.abs(y[0] - wl).max() <= TOL and np.abs(y[1] - wr).max() <= TOL
    assert np.abs(state - wstate).max() <= 1e-9


EQ_MODE_CASES = [
    ("parallel", dict(seed=7), dict(structure=1)),
    ("parallel_sat0_lr", dict(seed=8, modes=[i % 3 for i in range(20)]), dict(structure=1, saturation=0.0)),
    ("parallel_stress", dict(seed=7, stress=True), dict(structure=1)),
    ("agc", dict(seed=7), dict(agc=True)),
    ("agc_parallel", dict(seed=11), dict(agc=True, structure=1)),
    ("mid_side", dict(seed=7, modes=[0, 3, 4, 1, 2] * 4, flat=[5, 6, 7]), dict()),
    ("mid_side_all", dict(seed=12, modes=[3, 4] * 10), dict(saturation=0.0)),
    ("mid_side_agc", dict(seed=10, modes=[4, 0, 3, 0] * 5), dict(agc=True)),
    ("parallel_mid_side", dict(seed=7, modes=[0, 3, 4, 1, 2] * 4, flat=[5, 6]), dict(structure=1)),
    ("parallel_mid_side_only_agc", dict(seed=9, modes=[3, 4] * 10), dict(structure=1, agc=True)),
]


@pytest.mark.parametrize("name,bkw,kw", EQ_MODE_CASES)
@pytest.mark.parametrize("block,T", [(512, 512 * 50), (64, 64 * 300), (2048, 2048 * 13)])
def test_eq_modes_match_reference(checker, name, bkw, kw, block, T):
    """Parallel structure (Processing.cpp:1132-1228), AGC (:343-445), Mid/Side bands (:690-740)."""
    sr = 48000.0
    params = signals.band_params(**bkw)
    xl, xr = signals.log_sweep(T, sr)
    xl = 1.7 * xl + signals.noise(T, 5, 0.2)      # decorrelated, different levels: Mid != Side != 0, AGC has work to do
    xr = 0.4 * xr
    eng = ConvoPeqEngine(1, 2, sr, block, T)
    eng.set_eq(0, signals.to_band(params), kw.get("saturation", 0.2), 0.0, kw.get("structure", 0), kw.get("agc", False))
    y = np.stack([xl, xr]).copy()
    eng.process(y, capi.STAGE_EQ)
    agc = eng.agc_state(0) if kw.get("agc") else None
    eng.close()
    wl, wr, _ = checker.eq_run(signals.to_eqband(params), xl, xr, sr, block, **kw)
    scale = max(1.0, np.abs(wl).max(), np.abs(wr).max())
    assert np.abs(y[0] - wl).max() <= TOL * scale and np.abs(y[1] - wr).max() <= TOL * scale
    if agc is not None:
        assert 0.06 <= agc[2] <= 16.0 and agc[2] != 1.0


def test_eq_modes_batch_mixed_streams(checker):
    """Several streams with different structures / AGC / Mid-Side settings in one handle, conv -> EQ -> epilogue, odd chunking."""
    sr, block, T, ir_len = 48000.0, 512, 16384, 20000
    settings = [dict(), dict(structure=1), dict(agc=True), dict(agc=True, structure=1), dict(), dict(agc=True), dict(), dict(structure=1)]
    bkws = [dict(seed=40), dict(seed=41), dict(seed=42, modes=[3, 0, 4, 1, 2] * 4), dict(seed=43, modes=[0, 4, 3, 0] * 5), dict(seed=44, modes=[4] * 20),
            dict(seed=45), dict(seed=46, modes=[0, 0, 3] + [0] * 17), dict(seed=47, modes=[2, 3] * 10)]
    n = len(settings)
    # serial and Parallel streams, with and without Mid/Side bands and AGC, share one handle
    eng = ConvoPeqEngine(n, 2, sr, block, T, conv_boundary=capi.CONV_OUTER, workspace_bytes=3 * 2 * (T // block + 40) * 512 * 16 * 5)
    x = np.stack([signals.noise(T, 500 + i, 0.3) for i in range(2 * n)])
    irs = [signals.synth_ir(ir_len, 600 + i) for i in range(2 * n)]
    for s in range(n):
        for ch in range(2):
            eng.set_impulse(s, ch, irs[2 * s + ch], 1.0, None)
        eng.set_eq(s, signals.to_band(signals.band_params(**bkws[s])), 0.2, 0.0, settings[s].get("structure", 0), settings[s].get("agc", False))
    eng.set_epilogue(1.1, 0)
    y = x.copy()
    eng.process(y, capi.STAGE_ALL)
    eng.close()
    for s in range(n):
        want = checker.chain_run((irs[2 * s], irs[2 * s + 1]), signals.to_eqband(signals.band_params(**bkws[s])), x[2 * s:2 * s + 2], sr, block,
                                 None, makeup=1.1, **settings[s])
        assert np.abs(y[2 * s:2 * s + 2] - want).max() <= TOL, s


@pytest.mark.parametrize("block,n_cb", [(441, 3), (441, 40), (1000, 9), (100, 7)])
def test_eq_final_state_when_the_signal_ends_inside_a_thread_block(checker, block, n_cb):
    """The carried band state is the state after sample T - 1 also when T is not a multiple of the 32 samples a thread holds."""
    sr, T = 44100.0, block * n_cb
    params = signals.band_params(seed=21, types=[i % 5 for i in range(20)])
    xl, xr = signals.log_sweep(T, sr)
    eng = ConvoPeqEngine(1, 2, sr, block, T)
    eng.set_eq(0, signals.to_band(params))
    y = np.stack([xl, xr]).copy()
    eng.process(y, capi.STAGE_EQ)
    st = eng.eq_state(0)
    eng.close()
    wl, wr, wst = checker.eq_run(signals.to_eqband(params), xl, xr, sr, block)
    assert np.abs(y[0] - wl).max() <= TOL and np.abs(y[1] - wr).max() <= TOL
    assert np.abs(st - wst).max() <= 1e-9 * max(1.0, np.abs(wst).max())


def test_eq_mono_stream(checker):
    sr, block, T = 48000.0, 512, 40960
    params = signals.band_params(seed=12, modes=[i % 3 for i in range(20)])
    xl, _ = signals.log_sweep(T, sr)
    eng = ConvoPeqEngine(1, 1, sr, block, T)
    eng.set_eq(0, signals.to_band(params))
    y = xl[None, :].copy()
    eng.process(y, capi.STAGE_EQ)
    eng.close()
    wl, _, _ = checker.eq_run(signals.to_eqband(params), xl, None, sr, block)
    assert np.abs(y[0] - wl).max() <= TOL


def test_eq_total_gain_ramp(checker):
    sr, block, T = 48000.0, 512, 512 * 60
    params = signals.band_params(seed=7)
    xl, xr = signals.log_sweep(T, sr)
    eng = ConvoPeqEngine(1, 2, sr, block, T)
    eng.set_eq(0, signals.to_band(params), 0.2, 0.0)
    eng.schedule_total_gain(0, 20, -6.0)
    y = np.stack([xl, xr]).copy()
    eng.process(y, capi.STAGE_EQ)
    eng.close()
    wl, wr, _ = checker.eq_run(signals.to_eqband(params), xl, xr, sr, block, gain_change_db=-6.0, gain_change_at=20 * block)
    assert np.abs(y[0] - wl).max() <= TOL and np.abs(y[1] - wr).max() <= TOL


def test_full_chain_batch_of_streams(checker, oracle):
    """cfg4 in miniature: several streams, distinct IRs and EQs, conv -> EQ -> makeup + headroom."""
    sr, block, T, n_streams, ir_len = 48000.0, 512, 32768, 5, 131072
    eng = ConvoPeqEngine(n_streams, 2, sr, block, T, conv_boundary=capi.CONV_OUTER)
    x = np.stack([signals.noise(T, 100 + i) for i in range(2 * n_streams)])
    irs = [signals.synth_ir(ir_len, 200 + i) for i in range(2 * n_streams)]
    cspec, ospec = capi.default_filter_spec(), OFilterSpec()
    for s in range(n_streams):
        for ch in range(2):
            eng.set_impulse(s, ch, irs[2 * s + ch], 1.0, cspec)
        eng.set_eq(s, signals.to_band(signals.band_params(300 + s)))
    eng.set_epilogue(makeup_gain=1.3, dither_bits=0)
    y = x.copy()
    eng.process(y, capi.STAGE_ALL)
    t = eng.timings()
    assert t.kernel_launches >= 7
    eng.close()
    for s in range(n_streams):
        w = []
        for ch in range(2):
            c, _ = checker.nuc_run(irs[2 * s + ch], x[2 * s + ch], block, spec=ospec)
            w.append(oracle.outer_wet(c, 1.0))
        wl, wr, _ = checker.eq_run(signals.to_eqband(signals.band_params(300 + s)), w[0], w[1], sr, block)
        for ch, wv in enumerate((wl, wr)):
            want, _, _ = oracle.epilogue(wv, 1.3, sr, 0)
            assert np.abs(y[2 * s + ch] - want).max() <= TOL, (s, ch)


@pytest.mark.parametrize("sr,bits,nch,block,T", [(48000.0, 24, 2, 512, 65536), (96000.0, 16, 2, 480, 48000), (44100.0, 32, 1, 64, 8192),
                                                 (192000.0, 24, 2, 100, 20000)])
def test_dither_epilogue(checker, sr, bits, nch, block, T):
    """Injected uniforms (the VSL ring's replacement).  The shaper is chaotic, so the check is bit-for-bit: the quantised
    output and the carried error history equal the reference's own PsychoacousticDither.h compiled in place (checker = Ref;
    the restatement, itself pinned bit-for-bit against it, elsewhere).  70 sequences = three warps, one partly filled."""
    n_streams = 70 // nch
    rng = np.random.default_rng(5)
    x = np.stack([signals.noise(T, 100 + i, 0.3) for i in range(n_streams * nch)])
    u = rng.random((n_streams * nch, 2 * T))
    eng = ConvoPeqEngine(n_streams, nch, sr, block, T)
    eng.set_epilogue(0.9, bits, u)
    y = x.copy()
    eng.process(y, capi.STAGE_EPILOGUE)
    eng.close()
    for s in range(0, n_streams, 7):
        rows = slice(s * nch, (s + 1) * nch)
        want, _ = checker.dither_run(x[rows] * 0.9, u[rows], sr, bits, block)
        assert np.array_equal(y[rows], want), s


def test_dither_with_the_reference_fallback_generator(checker):
    """cpq_set_dither_seed: uniforms drawn on the device by the reference's own fallback generator (one PsychoacousticDither(seed)
    per stream); bit for bit against the reference compiled in place with its VSL stream failing."""
    sr, bits, block, T, n_streams = 48000.0, 24, 512, 512 * 50 + 512, 37
    x = np.stack([signals.noise(T, 300 + i, 0.3) for i in range(2 * n_streams)])
    seeds = [(0x9E3779B97F4A7C15 * (s + 1)) & 0xFFFFFFFFFFFFFFFF for s in range(n_streams)]
    eng = ConvoPeqEngine(n_streams, 2, sr, block, T)
    eng.set_epilogue(0.9, bits)
    eng.set_dither_seed(seeds)
    y = x.copy()
    eng.process(y, capi.STAGE_EPILOGUE)
    y2 = x.copy()
    eng.process(y2, capi.STAGE_EPILOGUE)        # one-shot mode: a fresh generator for every call
    eng.close()
    assert np.array_equal(y, y2)
    for s in range(0, n_streams, 6):
        want, _ = checker.dither_run_seeded(x[2 * s:2 * s + 2] * 0.9, seeds[s], sr, bits, block)
        assert np.array_equal(y[2 * s:2 * s + 2], want), s


@pytest.mark.parametrize("seeded", [False, True])
def test_dither_time_segments_equal_the_one_shot_call(seeded):
    """A device-resident conv -> EQ -> dither call runs in time segments (the shaper of one segment beside the transforms of the
    next, through the streaming continuation); the host-buffer call runs in one piece.  The shaper is chaotic, so equal means
    bit for bit -- everything up to the quantiser is identical because the segments end on EQ tile boundaries.  (Both calls use
    look-back links: at most 32 sequences per chunk.)"""
    import torch
    sr, block, T, n_streams = 48000.0, 512, 65536 * 3 + 512 * 3, 40
    n_seq = 2 * n_streams
    x = np.stack([signals.noise(T, 700 + i, 0.3) for i in range(n_seq)])
    eng = ConvoPeqEngine(n_streams, 2, sr, block, T, conv_boundary=capi.CONV_OUTER, workspace_bytes=100 << 20)   # several small sequence chunks
    for s in range(n_streams):
        for ch in range(2):
            eng.set_impulse(s, ch, signals.synth_ir(20000, 720 + (2 * s + ch) % 7), 1.0, capi.default_filter_spec())
        eng.set_eq(s, signals.to_band(signals.band_params(740 + s % 5)))
    if seeded:
        eng.set_epilogue(0.9, 24)
        eng.set_dither_seed([(0x9E3779B97F4A7C15 * (s + 1)) & 0xFFFFFFFFFFFFFFFF for s in range(n_streams)])
    else:
        eng.set_epilogue(0.9, 24, np.random.default_rng(6).random((n_seq, 2 * T)))
    eng.set_peak_limiter(100.0)
    y = x.copy()
    eng.process(y, capi.STAGE_ALL)                       # host buffers: one piece
    d = torch.from_numpy(x).cuda()
    eng.process_device(d.data_ptr(), T, T, capi.STAGE_ALL)   # device-resident: time segments
    t = eng.timings()
    d2 = torch.from_numpy(x).cuda()
    eng.process_device(d2.data_ptr(), T, T, capi.STAGE_ALL)  # and again: a new stream starts from Reset
    eng.close()
    import os
    assert t.chunks >= 3 and (os.environ.get("CPQ_DITHER_SEGMENTS") == "1" or t.chunks % 3 == 0)   # three segments, at most 32 sequences per chunk
    assert np.array_equal(d.cpu().numpy(), y) and np.array_equal(d2.cpu().numpy(), y)


def test_dither_with_a_block_that_does_not_divide_the_eq_tile_stays_in_one_piece():
    """480-sample callbacks: time segments would not be whole callbacks and whole EQ tiles at once, so the device-resident call
    runs in one piece like the host call -- same bits."""
    import torch
    sr, block, T, n_streams = 48000.0, 480, 480 * 300, 3
    x = np.stack([signals.noise(T, 760 + i, 0.3) for i in range(2 * n_streams)])
    eng = ConvoPeqEngine(n_streams, 2, sr, block, T)
    for s in range(n_streams):
        for ch in range(2):
            eng.set_impulse(s, ch, signals.synth_ir(20000, 770 + 2 * s + ch), 1.0, capi.default_filter_spec())
        eng.set_eq(s, signals.to_band(signals.band_params(780 + s)))
    eng.set_epilogue(0.9, 24)
    eng.set_dither_seed([11, 12, 13])
    y = x.copy()
    eng.process(y, capi.STAGE_ALL)
    d = torch.from_numpy(x).cuda()
    eng.process_device(d.data_ptr(), T, T, capi.STAGE_ALL)
    chunks = eng.timings().chunks
    eng.close()
    assert chunks == 1 and np.array_equal(d.cpu().numpy(), y)


def test_pageable_host_buffers_go_through_the_staging_threads():
    """A pageable caller buffer of more than 32 MB is staged through pinned slots by host threads (HostStager), chunk by chunk;
    smaller ones are left to the driver.  Same samples either way: the convolver is bit-identical whatever the chunking, so the
    staged call, the pinned call and the device-resident call must agree exactly -- rows at an irregular pitch included."""
    import torch
    sr, block, T, n_streams = 48000.0, 512, 512 * 200, 27      # 54 sequences x 102400 samples = 44 MB
    n_seq = 2 * n_streams
    eng = ConvoPeqEngine(n_streams, 2, sr, block, T)
    for s in range(n_streams):
        for ch in range(2):
            eng.set_impulse(s, ch, signals.synth_ir(9000, 30 + (2 * s + ch) % 5), 1.0, None)
    x = np.stack([signals.noise(T, 40 + i % 9, 0.3) * (1 + i) for i in range(n_seq)])
    staged = x.copy()
    eng.process(staged, capi.STAGE_CONV)                     # pageable, above the threshold: staging threads
    rows = [x[i].copy() for i in range(n_seq)]                # separately allocated rows: no common pitch
    ptrs = (C.POINTER(C.c_double) * n_seq)(*[r.ctypes.data_as(C.POINTER(C.c_double)) for r in rows])
    assert eng.lib.cpq_process(eng.h, ptrs, T, capi.STAGE_CONV) == capi.OK
    pinned = torch.from_numpy(x).pin_memory()
    eng.process_host_ptrs(pinned.data_ptr(), T, T, capi.STAGE_CONV)
    dev = torch.from_numpy(x).cuda()
    eng.process_device(dev.data_ptr(), T, T, capi.STAGE_CONV)
    small = x[:, :512 * 4].copy()
    eng.process(small, capi.STAGE_CONV)                      # 1.7 MB: the driver's staging
    eng.close()
    want = dev.cpu().numpy()
    assert np.abs(want).max() > 1e-3
    assert np.array_equal(staged, want) and np.array_equal(pinned.numpy(), want) and np.array_equal(np.stack(rows), want)
    assert np.array_equal(small, want[:, :512 * 4])


def test_partition_range_partials_sum_to_full(checker):
    """cfg5 in miniature: the convolver is linear in the IR, so rank partials over partition ranges add up."""
    ir_len, block, T = 131072, 512, 32768
    ir = signals.synth_ir(ir_len, 9)
    x = np.stack([signals.noise(T, 1), signals.noise(T, 2)])
    eng = ConvoPeqEngine(1, 2, 48000.0, block, T)
    eng.set_impulse(0, 0, ir)
    eng.set_impulse(0, 1, ir)
    total = eng.total_partitions()
    acc = np.zeros_like(x)
    ranks = 4
    for r in range(ranks):
        b, e = total * r // ranks, total * (r + 1) // ranks
        eng.set_partition_range(b, e)
        y = x.copy()
        eng.process(y, capi.STAGE_CONV)
        acc += y
    eng.close()
    want, _ = checker.nuc_run(ir, x[0], block)
    assert np.abs(acc[0] - want).max() <= TOL


@pytest.mark.parametrize("agc", [False, True])
def test_partials_summed_in_the_eq_load(checker, agc):
    """SURVEY 8e, fused reduce: three ranks' convolver partials sit in three buffers (here on one GPU); every 'rank' finishes
    only its stream window, summing the buffers in rank order while its EQ launch loads the tiles (cpq_set_partial_sources +
    cpq_set_stream_window).  The result equals the reference chain, and no rank ever forms the full sum in a separate pass."""
    import torch
    sr, block, T, ir_len, n_streams, ranks = 48000.0, 512, 32768, 100000, 3, 3
    irs = [signals.synth_ir(ir_len, 20 + ch) for ch in range(2)]
    x = np.stack([signals.noise(T, 30 + i) for i in range(2 * n_streams)])
    bkw = [dict(seed=50), dict(seed=51, modes=[0, 3, 4, 0] * 5), dict(seed=52)]
    # (inner boundary: on one handle the deferred outer step of a partial convolution is armed once per call pair, which
    # this single-process emulation of three ranks would consume with the first window)
    eng = ConvoPeqEngine(n_streams, 2, sr, block, T, shared_ir=True)
    for ch in range(2):
        eng.set_impulse(-1, ch, irs[ch], 1.0, capi.default_filter_spec())
    for st in range(n_streams):
        eng.set_eq(st, signals.to_band(signals.band_params(**bkw[st])), agc=agc)
    eng.set_epilogue(1.2, 0)
    total = eng.total_partitions()
    bufs = [torch.from_numpy(x).cuda() for _ in range(ranks)]
    for r in range(ranks):                                   # every rank: its partition range, all streams
        eng.set_partition_range(total * r // ranks, total * (r + 1) // ranks)
        eng.process_device(bufs[r].data_ptr(), T, T, capi.STAGE_CONV)
    eng.set_partition_range(0, -1)
    eng.set_partial_sources([b.data_ptr() for b in bufs])
    for r in range(ranks):                                   # every rank: finish the stream it owns, reading all partials
        eng.set_stream_window(r, 1)
        eng.process_device(bufs[r].data_ptr(), T, T, capi.STAGE_EQ | capi.STAGE_EPILOGUE)
    eng.set_partial_sources([])
    eng.set_stream_window(0, -1)
    torch.cuda.synchronize()
    eng.close()
    for st in range(n_streams):
        got = bufs[st][2 * st:2 * st + 2].cpu().numpy()
        want = checker.chain_run(irs, signals.to_eqband(signals.band_params(**bkw[st])), x[2 * st:2 * st + 2], sr, block, OFilterSpec(),
                                 makeup=1.2, agc=agc, outer=False)
        assert np.abs(got - want).max() <= TOL, st


# ---- golden vectors produced by the reference's own code (tests/golden/make_golden.py) ----
import os as _os
from tests.golden.cases import CONV_CASES as _GCONV, EQ_CASES as _GEQ, CHAIN_CASES as _GCHAIN, conv_inputs, eq_inputs, chain_inputs

_GOLD = np.load(_os.path.join(_os.path.dirname(__file__), "golden", "golden.npz"))


@pytest.mark.parametrize("name", sorted(_GCONV))
def test_convolver_matches_golden(name):
    # (direct_head cases set the handle flag before the impulses)
    c = _GCONV[name]
    ir, x = conv_inputs(c)
    cspec = capi.default_filter_spec(**c["spec"]) if c["spec"] is not None else None
    eng = ConvoPeqEngine(1, 1, 48000.0, c["block"], c["T"])
    eng.set_direct_head(c.get("direct_head", False))
    eng.set_impulse(0, 0, ir, c.get("scale", 1.0), cspec)
    y = x[None, :].copy()
    eng.process(y, capi.STAGE_CONV)
    lay = eng.layout()
    eng.close()
    assert np.abs(y[0] - _GOLD["conv/" + name]).max() <= TOL
    want = _GOLD["conv_layout/" + name]
    got = np.array([[lay.layers[i].part_size, lay.layers[i].num_parts_ir, lay.layers[i].parts_per_callback,
                     lay.layers[i].output_delay_samples] for i in range(lay.num_layers)])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", sorted(_GEQ))
def test_eq_matches_golden(name):
    c = _GEQ[name]
    bands, xl, xr = eq_inputs(c)
    kw = c.get("kw", {})
    eng = ConvoPeqEngine(1, 2, c["sr"], c["block"], c["T"])
    eng.set_eq(0, signals.to_band(bands), kw.get("saturation", 0.2), kw.get("total_gain_db", 0.0), kw.get("structure", 0),
               kw.get("agc", False))
    if "gain_change_db" in kw:
        eng.schedule_total_gain(0, kw["gain_change_at"] // c["block"], kw["gain_change_db"])
    y = np.stack([xl, xr]).copy()
    eng.process(y, capi.STAGE_EQ)
    eng.close()
    g = _GOLD["eq/" + name]
    assert np.abs(y - g).max() <= TOL * max(1.0, np.abs(g).max())


@pytest.mark.parametrize("name", sorted(_GCHAIN))
def test_chain_matches_golden(name):
    c = _GCHAIN[name]
    irs, bands, x = chain_inputs(c)
    eng = ConvoPeqEngine(1, 2, c["sr"], c["block"], c["T"], conv_boundary=capi.CONV_OUTER)
    for ch in range(2):
        eng.set_impulse(0, ch, irs[ch], 1.0, capi.default_filter_spec(**c["spec"]))
    eng.set_eq(0, signals.to_band(bands))
    eng.set_epilogue(c["makeup"], 0)
    y = x.copy()
    eng.process(y, capi.STAGE_ALL)
    eng.close()
    assert np.abs(y - _GOLD["chain/" + name]).max() <= TOL


def test_error_paths_fail_loudly():
    eng = ConvoPeqEngine(1, 2, 48000.0, 512, 4096)
    x = np.zeros((2, 4096))
    with pytest.raises(capi.CpqError) as e:
        eng.process(x, capi.STAGE_CONV)            # no impulse set
    assert e.value.status == capi.ERR_NOT_READY
    with pytest.raises(capi.CpqError):
        eng.process(np.zeros((2, 1000)), capi.STAGE_EQ)   # T not a multiple of the block
    with pytest.raises(capi.CpqError) as e:
        ConvoPeqEngine(1, 2, 44100.0, 32, 4096)          # host blocks below 64 (or above 8192) are outside the path
    assert e.value.status == capi.ERR_UNSUPPORTED
    bands = signals.to_band(signals.band_params(1, modes=[3] * 20))
    mono = ConvoPeqEngine(1, 1, 48000.0, 512, 4096)
    with pytest.raises(capi.CpqError) as e:
        mono.set_eq(0, bands)                      # Mid/Side bands need both channels of a stream
    assert e.value.status == capi.ERR_UNSUPPORTED
    mono.close()
    eng.set_eq(0, bands, agc=True)
    eng.schedule_total_gain(0, 2, -3.0)            # with AGC the reference never applies the total-gain ramp: refused
    with pytest.raises(capi.CpqError) as e:
        eng.process(np.zeros((2, 4096)), capi.STAGE_EQ)
    assert e.value.status == capi.ERR_UNSUPPORTED
    eng.set_impulse(0, 0, signals.synth_ir(4096, 1))
    with pytest.raises(capi.CpqError) as e:
        eng.set_impulse(0, 1, signals.synth_ir(100000, 1))   # different layer geometry in one handle
    assert e.value.status == capi.ERR_GEOMETRY
    eng.close()


def _poisoned(T, seed, kind):
    """A sweep + noise with samples the reference's state guard reacts to (Processing.cpp:174-175, :257-258)."""
    xl, xr = signals.log_sweep(T, 48000.0)
    xl = xl + signals.noise(T, seed, 0.1)
    xr = 0.5 * xr + signals.noise(T, seed + 1, 0.1)
    if kind == "nan":        # non-finite samples: output 0, both states of the first band zeroed
        xl[100], xr[5000], xl[8191], xl[8192], xr[20000] = np.nan, np.inf, -np.inf, np.nan, np.nan
        xl[T - 3] = np.inf
    elif kind == "huge":     # finite overflow: a state passes 1e15 and is zeroed, the other one may survive
        xl[100], xr[777], xl[9000], xr[9001] = 1e200, -3e16, 5e15, 1e15
        xl[T - 700] = 2e17
    elif kind == "dense":    # one event in every 1024-sample segment of the first tiles, several in some
        for k in range(0, min(T, 30000), 700):
            (xl if k % 1400 == 0 else xr)[k] = np.nan if k % 2100 else 1e180
    return xl, xr


@pytest.mark.parametrize("kind", ["nan", "huge", "dense"])
@pytest.mark.parametrize("structure", [0, 1])
@pytest.mark.parametrize("n_streams", [1, 20])   # look-back links (few sequences) and chained links
def test_eq_state_reset_matches_the_reference(checker, kind, structure, n_streams):
    """Where the reference zeroes a runaway state and carries on, so does the engine: the affected 1024-sample segment of the
    band runs serially with the literal recurrence and hands the true state on (serialBand); everything else stays on the
    blocked scan.  Every stream gets its events at different positions."""
    sr, block, T = 48000.0, 512, 8192 * 4 + 512 * 3
    eng = ConvoPeqEngine(n_streams, 2, sr, block, T)
    xs, ps = [], []
    for s in range(n_streams):
        p = signals.band_params(60 + s, stress=(s % 3 == 1), modes=[i % 3 for i in range(20)] if s % 4 == 2 else None)
        xl, xr = _poisoned(T, 70 + 2 * s, kind)
        xl, xr = np.roll(xl, 37 * s), np.roll(xr, 91 * s)
        eng.set_eq(s, signals.to_band(p), 0.2, 0.0, structure, False)
        xs += [xl, xr]
        ps.append(p)
    y = np.stack(xs).copy()
    eng.process(y, capi.STAGE_EQ)
    state = [eng.eq_state(s) for s in range(n_streams)]
    eng.close()
    for s in range(n_streams):
        wl, wr, wst = checker.eq_run(signals.to_eqband(ps[s]), xs[2 * s], xs[2 * s + 1], sr, block, structure=structure)
        for got, want in ((y[2 * s], wl), (y[2 * s + 1], wr)):
            # the Parallel structure adds the band differences to the *input*: a non-finite sample stays non-finite there
            fin = np.isfinite(want)
            assert (np.isfinite(got) == fin).all() and (structure == 1 or fin.all()), s
            # after a 1e200 sample a surviving state sits just below 1e15 and takes ten thousand samples to decay; below 1e9 the
            # blocked scan takes over again, with an absolute rounding of state x 1e-16: 1e-7 there
            assert np.abs(got[fin] - want[fin]).max() <= (1e-7 if kind == "huge" else 1e-9), s
        assert np.abs(np.asarray(state[s]).reshape(2, 20, 2) - wst).max() <= 1e-7 * max(1.0, np.abs(wst).max()), s


def test_eq_state_reset_through_the_whole_chain(checker):
    """A NaN input sample with the convolver at the inner boundary (no wet scrub): the first EQ band sees NaN for as long as a
    partition holds the poisoned frame -- thousands of consecutive resets; chain through cpq_process with the epilogue."""
    sr, block, T, ir_len = 48000.0, 512, 8192 * 3, 3000
    n = 40
    eng = ConvoPeqEngine(n, 2, sr, block, T, conv_boundary=capi.CONV_INNER)
    x = np.stack([signals.noise(T, 800 + i, 0.3) for i in range(2 * n)])
    irs = [signals.synth_ir(ir_len, 900 + i) for i in range(2 * n)]
    for s in range(n):
        x[2 * s, 1000 + 211 * s] = np.nan     # every frame it touches comes out of the convolver as NaN, in both implementations
        for ch in range(2):
            eng.set_impulse(s, ch, irs[2 * s + ch], 1.0, None)
        eng.set_eq(s, signals.to_band(signals.band_params(950 + s)))
    eng.set_epilogue(1.0, 0)
    y = x.copy()
    eng.process(y, capi.STAGE_ALL)
    eng.close()
    for s in range(0, n, 7):
        want = checker.chain_run((irs[2 * s], irs[2 * s + 1]), signals.to_eqband(signals.band_params(950 + s)), x[2 * s:2 * s + 2], sr, block,
                                 None, outer=False)
        assert np.abs(y[2 * s:2 * s + 2] - want).max() <= 1e-9, s


@pytest.mark.parametrize("modes,mono", [(None, False), ([i % 3 for i in range(20)], False), (None, True)])
def test_eq_large_signal_takes_the_exact_path(checker, modes, mono):
    """|out| >= 4.5 before saturation: the fast pass must hand over to the exact per-sample semantics (tanh clamp,
    +-100 clamp).  Beyond the threshold the reference's two band functions differ: processBandStereo (Stereo bands of a stereo
    stream) clamps the tanh argument, the scalar processBand (Left / Right bands, mono streams) returns +-1 -- both mirrored."""
    sr, block, T = 48000.0, 512, 8192
    params = signals.band_params(7, stress=True, modes=modes)
    xl, xr = signals.log_sweep(T, sr, amp=30.0)
    eng = ConvoPeqEngine(1, 1 if mono else 2, sr, block, T)
    eng.set_eq(0, signals.to_band(params))
    y = (xl[None, :] if mono else np.stack([xl, xr])).copy()
    eng.process(y, capi.STAGE_EQ)
    eng.close()
    wl, wr, _ = checker.eq_run(signals.to_eqband(params), xl, None if mono else xr, sr, block)
    assert np.abs(wl).max() > 4.5
    assert np.abs(y[0] - wl).max() <= 1e-9 and (mono or np.abs(y[1] - wr).max() <= 1e-9)


def test_shared_ir_and_eq_across_sequence_chunks(checker):
    """One IR pair / one EQ shared by all streams; the host entry point splits the batch into sequence chunks,
    so channel -> IR row mapping must use absolute sequence indices."""
    sr, block, T, n_streams = 48000.0, 512, 8192, 5
    irs = [signals.synth_ir(9000, 40), signals.synth_ir(9000, 41)]
    x = np.stack([signals.noise(T, 700 + i) for i in range(2 * n_streams)])
    eng = ConvoPeqEngine(n_streams, 2, sr, block, T, shared_ir=True, shared_eq=True)
    for ch in range(2):
        eng.set_impulse(-1, ch, irs[ch])
    params = signals.band_params(5)
    eng.set_eq(-1, signals.to_band(params))
    y = x.copy()
    eng.process(y, capi.STAGE_CONV | capi.STAGE_EQ)
    eng.close()
    for s in range(n_streams):
        c = [checker.nuc_run(irs[ch], x[2 * s + ch], block)[0] for ch in range(2)]
        wl, wr, _ = checker.eq_run(signals.to_eqband(params), c[0], c[1], sr, block)
        assert np.abs(y[2 * s] - wl).max() <= TOL and np.abs(y[2 * s + 1] - wr).max() <= TOL


# ---- output stages: OutputFilter -> makeup -> DC blocker -> headroom -> scrub + clamp ----
from tests.golden.cases import OUTPUT_CASES as _GOUT, FULL_CHAIN_CASES as _GFULL, output_inputs


def _gpu_output(x, sr, block, stages, conv_is_last=False, hc=1, lc=0, lp=1, makeup=1.0, dc_cutoff=3.0, clamp=True, use_filter=True,
                n_channels=2, limiter_ms=0.0):
    T = x.shape[1]
    eng = ConvoPeqEngine(x.shape[0] // n_channels, n_channels, sr, block, T)
    eng.set_output_filter(use_filter, conv_is_last, hc, lc, lp)
    eng.set_output_stage(dc_cutoff, clamp)
    eng.set_peak_limiter(limiter_ms)
    eng.set_epilogue(makeup, 0)
    y = x.copy()
    eng.process(y, stages)
    eng.close()
    return y


@pytest.mark.parametrize("name", sorted(_GOUT))
def test_output_stage_matches_golden(name):
    c = _GOUT[name]
    y = _gpu_output(output_inputs(c), c["sr"], c["block"], capi.STAGE_OUTPUT_FILTER | capi.STAGE_EPILOGUE, **c["kw"])
    assert np.abs(y - _GOLD["output/" + name]).max() <= TOL
    d = np.zeros(15)
    kw = c["kw"]
    capi.load().cpq_output_filter_design(c["sr"], int(kw.get("conv_is_last", False)), kw.get("hc", 1), kw.get("lc", 0), kw.get("lp", 1),
                                         d.ctypes.data_as(capi.C.POINTER(capi.C.c_double)))
    assert np.array_equal(d.reshape(3, 5), _GOLD["output_design/" + name])   # same libm sin/cos as OutputFilter::prepare


@pytest.mark.parametrize("kw", [dict(), dict(conv_is_last=True, hc=0, lc=1), dict(conv_is_last=True, hc=2, makeup=1.7),
                                dict(lp=0, dc_cutoff=0.0), dict(use_filter=False), dict(clamp=False)])
@pytest.mark.parametrize("T,block,sr", [(8192 * 5 + 512, 512, 48000.0), (96000, 64, 192000.0)])
def test_output_stage_matches_reference_across_tiles(checker, kw, T, block, sr):
    """Several 8192-sample tiles (tile-to-tile chain of the output stages), a partial last tile, mono and stereo streams."""
    x = np.stack([signals.noise(T, 700 + i) for i in range(4)]) * 2.0 + 0.02
    y = _gpu_output(x, sr, block, capi.STAGE_OUTPUT_FILTER | capi.STAGE_EPILOGUE, **kw)
    for s in range(2):
        want = checker.output_run(x[2 * s:2 * s + 2], sr, block, **kw)
        assert np.abs(y[2 * s:2 * s + 2] - want).max() <= TOL
    ym = _gpu_output(x[:1], sr, block, capi.STAGE_OUTPUT_FILTER | capi.STAGE_EPILOGUE, n_channels=1, **kw)
    assert np.abs(ym - checker.output_run(x[:1], sr, block, **kw)).max() <= TOL


@pytest.mark.parametrize("kw", [dict(limiter_ms=100.0), dict(limiter_ms=100.0, use_filter=False, dc_cutoff=0.0), dict(limiter_ms=30.0, clamp=False),
                                dict(limiter_ms=100.0, makeup=0.05)])
def test_peak_limiter_matches_reference(checker, kw):
    """SimplePeakLimiter (audioengine/SimplePeakLimiter.h) between the scrub and the hard clamp: three stereo streams of which
    one stays below the knee for the whole signal (the stage is then exactly the identity and its serial kernel skips the
    stream), one is loud throughout and one has a loud burst followed by the release tail."""
    sr, block, T = 48000.0, 512, 8192 * 4 + 512
    x = np.stack([signals.noise(T, 800 + i) for i in range(6)])
    x[0:2] *= 0.5                      # quiet
    x[2:4] *= 6.0                      # loud
    x[4:6] *= 0.5
    x[4:6, 9000:9600] *= 14.0          # burst, then release
    y = _gpu_output(x, sr, block, capi.STAGE_OUTPUT_FILTER | capi.STAGE_EPILOGUE, **kw)
    for s in range(3):
        want = checker.output_run(x[2 * s:2 * s + 2], sr, block, **kw)
        assert np.abs(y[2 * s:2 * s + 2] - want).max() <= TOL, s
    if kw.get("makeup", 1.0) == 1.0:
        off = checker.output_run(x[2:4], sr, block, **{**kw, "limiter_ms": 0.0})
        assert np.abs(y[2:4] - off).max() > 1e-2          # the limiter acted on the loud stream
    ym = _gpu_output(x[2:3], sr, block, capi.STAGE_OUTPUT_FILTER | capi.STAGE_EPILOGUE, n_channels=1, **kw)
    assert np.abs(ym - checker.output_run(x[2:3], sr, block, **kw)).max() <= TOL


@pytest.mark.parametrize("name", sorted(_GFULL))
def test_full_chain_matches_golden(name):
    """conv -> wet gain -> EQ -> total gain -> OutputFilter -> makeup -> DC blocker -> headroom -> scrub + clamp in one call."""
    c = _GFULL[name]
    irs, bands, x = chain_inputs(c)
    eng = ConvoPeqEngine(1, 2, c["sr"], c["block"], c["T"], conv_boundary=capi.CONV_OUTER)
    for ch in range(2):
        eng.set_impulse(0, ch, irs[ch], 1.0, capi.default_filter_spec(**c["spec"]))
    eng.set_eq(0, signals.to_band(bands))
    eng.set_epilogue(c["makeup"], 0)
    eng.set_output_filter(True, c["out"]["conv_is_last"], 1, 0, c["out"]["lp"])
    eng.set_output_stage(c["out"]["dc_cutoff"], True)
    y = x.copy()
    eng.process(y, capi.STAGE_FULL)
    eng.close()
    assert np.abs(y - _GOLD["full_chain/" + name]).max() <= TOL


def test_output_filter_after_total_gain_ramp(checker):
    """The total-gain ramp sits between the EQ bands and OutputFilter: it is applied in registers before the output stages."""
    sr, block, T = 48000.0, 512, 512 * 60
    params = signals.band_params(seed=9)
    xl, xr = signals.log_sweep(T, sr)
    eng = ConvoPeqEngine(1, 2, sr, block, T)
    eng.set_eq(0, signals.to_band(params), 0.2, 0.0)
    eng.schedule_total_gain(0, 20, -6.0)
    eng.set_output_filter(True, False, 1, 0, 2)
    eng.set_output_stage(3.0, True)
    eng.set_epilogue(1.1, 0)
    y = np.stack([xl, xr]).copy()
    eng.process(y, capi.STAGE_EQ | capi.STAGE_OUTPUT_FILTER | capi.STAGE_EPILOGUE)
    eng.close()
    wl, wr, _ = checker.eq_run(signals.to_eqband(params), xl, xr, sr, block, gain_change_db=-6.0, gain_change_at=20 * block)
    want = checker.output_run(np.stack([wl, wr]), sr, block, lp=2, makeup=1.1)
    assert np.abs(y - want).max() <= TOL


@pytest.mark.parametrize("trim", [1.0, 0.7079457843841379])
def test_eq_then_convolver_order(checker, oracle, trim):
    """ProcessingOrder::EQThenConvolver: EQ -> input trim -> convolver (outer boundary) -> OutputFilter(convIsLast) -> epilogue."""
    sr, block, T, ir_len = 48000.0, 512, 24576, 65536
    x = np.stack([signals.noise(T, 810), signals.noise(T, 811)])
    irs = [signals.synth_ir(ir_len, 820), signals.synth_ir(ir_len, 821)]
    params = signals.band_params(830)
    cspec, ospec = capi.default_filter_spec(), OFilterSpec()
    eng = ConvoPeqEngine(1, 2, sr, block, T, conv_boundary=capi.CONV_OUTER)
    for ch in range(2):
        eng.set_impulse(0, ch, irs[ch], 1.0, cspec)
    eng.set_eq(0, signals.to_band(params))
    eng.set_conv_input_trim(trim)
    eng.set_output_filter(True, True, 1, 0, 1)
    eng.set_output_stage(3.0, True)
    eng.set_epilogue(1.2, 0)
    y = x.copy()
    eng.process(y, capi.STAGE_FULL | capi.ORDER_EQ_THEN_CONV)
    eng.close()
    wl, wr, _ = checker.eq_run(signals.to_eqband(params), x[0], x[1], sr, block)
    w = []
    for ch, v in enumerate((wl, wr)):
        c, _ = checker.nuc_run(irs[ch], v * trim, block, spec=ospec)
        w.append(oracle.outer_wet(c, 1.0))
    want = checker.output_run(np.stack(w), sr, block, conv_is_last=True, makeup=1.2)
    assert np.abs(y - want).max() <= TOL


@pytest.mark.parametrize("order", ["conv_eq", "eq_conv"])
def test_everything_at_once(checker, oracle, order):
    """All the optional stages of the path in one call, both processing orders, three streams with different settings:
    Mid/Side + AGC EQ, direct-form head, dry/wet mix with latency compensation, input trim, OutputFilter, DC blocker, peak limiter,
    clamp -- against the same chain assembled from the reference pieces."""
    from convopeq_b200.engine import ir_peak_latency
    sr, block, T, ir_len, n = 48000.0, 512, 8192 * 3, 50000, 3
    mix, trim, makeup = 0.6, 0.8, 2.5
    x = np.stack([signals.noise(T, 900 + i, 0.4) for i in range(2 * n)])
    irs = [np.roll(signals.synth_ir(ir_len, 920 + i), 200) for i in range(2 * n)]
    bkws = [dict(seed=940, modes=[0, 3, 4, 1, 2] * 4), dict(seed=941), dict(seed=942, modes=[4, 3] * 10)]
    ekw = [dict(agc=True), dict(structure=1), dict()]
    cspec, ospec = capi.default_filter_spec(), OFilterSpec()
    eng = ConvoPeqEngine(n, 2, sr, block, T, conv_boundary=capi.CONV_OUTER, workspace_bytes=40 << 20)
    eng.set_direct_head(True)
    for s in range(n):
        for ch in range(2):
            eng.set_impulse(s, ch, irs[2 * s + ch], 0.9, cspec)
        eng.set_eq(s, signals.to_band(signals.band_params(**bkws[s])), 0.2, 0.0, ekw[s].get("structure", 0), ekw[s].get("agc", False))
    delay = max(ir_peak_latency(irs[2 * s], irs[2 * s + 1]) for s in range(n))     # direct head: algorithm latency 0
    eng.set_mix(mix, delay)
    eng.set_conv_input_trim(trim)
    eng.set_output_filter(True, order == "eq_conv", 1, 0, 1)
    eng.set_output_stage(3.0, True)
    eng.set_peak_limiter(100.0)
    eng.set_epilogue(makeup, 0)
    y = x.copy()
    eng.process(y, capi.STAGE_FULL | (capi.ORDER_EQ_THEN_CONV if order == "eq_conv" else 0))
    eng.close()

    def conv(v, s):
        out = []
        for ch in range(2):
            wet, _ = checker.nuc_run(irs[2 * s + ch], v[ch], block, scale=0.9, spec=ospec, direct_head=True)
            out.append(oracle.outer_mix(wet, v[ch], mix, delay))
        return np.stack(out)

    def eq(v, s):
        l, r, _ = checker.eq_run(signals.to_eqband(signals.band_params(**bkws[s])), v[0], v[1], sr, block, **ekw[s])
        return np.stack([l, r])

    for s in range(n):
        v = x[2 * s:2 * s + 2]
        mid = eq(conv(v, s), s) if order == "conv_eq" else conv(eq(v, s) * trim, s)
        want = checker.output_run(mid, sr, block, conv_is_last=(order == "eq_conv"), makeup=makeup, limiter_ms=100.0)
        assert np.abs(y[2 * s:2 * s + 2] - want).max() <= TOL, (order, s)
        assert np.abs(want).max() > 0.5


@pytest.mark.parametrize("block,gain", [(512, 1.0), (441, 0.5), (64, 1.0 + 1e-10)])
def test_input_stage_matches_reference(checker, block, gain):
    """DSPCore::processInput's transform (InputBitDepthTransform.h compiled from the reference tree): gain, NaN / denormal scrub,
    clamp to [-1, 1]; +-Inf clamps in the four-wide body and is zeroed in the scalar remainder of a callback."""
    T = block * 20
    x = np.stack([signals.noise(T, 70 + i, 0.6) for i in range(4)])
    x[0, 3], x[0, 10], x[1, 11], x[2, 20], x[3, 21], x[1, 30] = np.nan, np.inf, -np.inf, 1e-25, -3e-21, 2.5
    x[2, block - 1], x[3, 2 * block - 1] = np.inf, np.nan          # last sample of a callback: scalar remainder when block % 4 != 0
    eng = ConvoPeqEngine(2, 2, 48000.0, block, T)
    eng.set_input_gain(gain)
    y = x.copy()
    eng.process(y, capi.STAGE_INPUT)
    # followed by the EQ in one call
    eng.set_eq(0, signals.to_band(signals.band_params(75)))
    eng.set_eq(1, signals.to_band(signals.band_params(76)))
    z = x.copy()
    eng.process(z, capi.STAGE_INPUT | capi.STAGE_EQ)
    eng.close()
    want = np.stack([np.concatenate([checker.input_transform(row[c * block:(c + 1) * block], gain) for c in range(20)]) for row in x])
    assert np.array_equal(y, want)
    for s in range(2):
        l, r, _ = checker.eq_run(signals.to_eqband(signals.band_params(75 + s)), want[2 * s], want[2 * s + 1], 48000.0, block)
        assert np.abs(z[2 * s] - l).max() <= TOL and np.abs(z[2 * s + 1] - r).max() <= TOL


@pytest.mark.parametrize("n_streams,T", [(1, 8192), (45, 512 * 24)])
def test_float_host_buffers(n_streams, T):
    """cpq_process_f32: FP32 on the wire, FP64 arithmetic -- bit-identical to the double entry point fed the same (float-valued)
    samples and rounded to float afterwards; 90 sequences move in ~30 chunks through the 3-in / 2-out staging slots."""
    sr, block, ir_len = 48000.0, 512, 20000
    eng = ConvoPeqEngine(n_streams, 2, sr, block, T, conv_boundary=capi.CONV_OUTER, shared_ir=True)
    for ch in range(2):
        eng.set_impulse(-1, ch, signals.synth_ir(ir_len, 30 + ch), 1.0, capi.default_filter_spec())
    for s in range(n_streams):
        eng.set_eq(s, signals.to_band(signals.band_params(100 + s % 5)))
    eng.set_epilogue(1.0, 0)
    xf = np.stack([signals.noise(T, 200 + i, 0.3) for i in range(2 * n_streams)]).astype(np.float32)
    y64 = xf.astype(np.float64)
    eng.process(y64, capi.STAGE_INPUT | capi.STAGE_ALL)
    y32 = xf.copy()
    eng.process_f32(y32, capi.STAGE_INPUT | capi.STAGE_ALL)
    eng.close()
    assert np.array_equal(y32, y64.astype(np.float32))
    assert np.abs(y32).max() > 1e-3
