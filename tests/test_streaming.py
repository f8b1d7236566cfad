"""Streaming continuation (cpq_set_streaming): the state the reference keeps between callbacks -- FDL, partial input frame,
delay-line cursors (MKLNonUniformConvolver.h:288-365, .cpp:1407-1548, cleared only by Reset :1693), EQ filterState
(EQProcessor.h:637), AGC / limiter envelopes, dither error history -- carried across cpq_process calls.  A signal processed in
segments of 1, 7 and 64 callbacks must equal the one-shot call (convolver: bit for bit; EQ: to rounding, its scan tiles start
at the segment boundary) and the reference (<= 1e-10)."""
import numpy as np
import pytest

from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
from oracle.bindings import FilterSpec as OFilterSpec
from tests import signals

pytestmark = pytest.mark.gpu
TOL = 1e-10

CONFIGS = {
    # name: (sample rate, IR taps, callbacks)   block 512
    "cfg1a": (48000.0, 65536, 200),      # L0 12x512 + L1 15x4096, D1 7168
    "cfg3": (96000.0, 262144, 200),      # L0 23x512 + L1 62x4096
    "three_layers_b64": (48000.0, 65536, 0),   # block 64: 64 / 512 / 4096, filled in below
}


def _engine(sr, ir_len, T, block, n_streams=2, spec_kw=None):
    eng = ConvoPeqEngine(n_streams, 2, sr, block, T, conv_boundary=capi.CONV_OUTER)
    spec = capi.default_filter_spec(sample_rate=sr, **(spec_kw or {}))
    irs = [signals.synth_ir(ir_len, 70 + i) for i in range(2 * n_streams)]
    for s in range(n_streams):
        for ch in range(2):
            eng.set_impulse(s, ch, irs[2 * s + ch], 1.0, spec)
        eng.set_eq(s, signals.to_band(signals.band_params(80 + s)))
    eng.set_epilogue(1.1, 0)
    return eng, irs


def _segmented(eng, x, seg_samples, stages):
    y = np.empty_like(x)
    T = x.shape[1]
    for t0 in range(0, T, seg_samples):
        t1 = min(T, t0 + seg_samples)
        part = np.ascontiguousarray(x[:, t0:t1])
        eng.process(part, stages)
        y[:, t0:t1] = part
    return y


@pytest.mark.parametrize("seg", [1, 7, 64])
@pytest.mark.parametrize("name", ["cfg1a", "cfg3"])
def test_segments_equal_one_shot_and_reference(checker, name, seg):
    sr, ir_len, n_cb = CONFIGS[name]
    block = 512
    T = n_cb * block
    eng, irs = _engine(sr, ir_len, T, block)
    x = np.stack([signals.noise(T, 90 + i) for i in range(4)])
    # one-shot
    conv1 = x.copy()
    eng.process(conv1, capi.STAGE_CONV)
    full1 = x.copy()
    eng.process(full1, capi.STAGE_ALL)
    # segmented, convolver alone: identical bits
    eng.set_streaming(True)
    conv2 = _segmented(eng, x, seg * block, capi.STAGE_CONV)
    assert eng.stream_position() == T
    assert np.array_equal(conv1, conv2), np.abs(conv1 - conv2).max()
    # segmented, whole chain
    eng.reset()
    full2 = _segmented(eng, x, seg * block, capi.STAGE_ALL)
    state2 = eng.eq_state(0)
    eng.close()
    assert np.abs(full1 - full2).max() <= 1e-12
    for s in range(2):
        want = checker.chain_run((irs[2 * s], irs[2 * s + 1]), signals.to_eqband(signals.band_params(80 + s)), x[2 * s:2 * s + 2], sr, block,
                                 OFilterSpec(sample_rate=sr), makeup=1.1)
        assert np.abs(full2[2 * s:2 * s + 2] - want).max() <= TOL, s
    assert np.isfinite(state2).all()


@pytest.mark.parametrize("block,seg", [(480, 1), (480, 7), (480, 64), (441, 1), (441, 5), (1000, 3), (96, 11)])
def test_host_blocks_that_are_not_a_power_of_two(checker, block, seg):
    """480- / 441-sample hosts calling per callback (or a few): layer 0 then goes through the reference's output ring, whose
    un-read samples are carried like a tail's delay line (Get's varying delivery per callback included).  441 x an odd number
    of callbacks makes the calls odd-sized as well."""
    sr, ir_len, n_cb = 48000.0, 65536, 130
    T = n_cb * block
    known = 64
    while known < block:
        known *= 2
    eng, irs = _engine(sr, ir_len, T, block)
    x = np.stack([signals.noise(T, 390 + i) for i in range(4)])
    conv1 = x.copy()
    eng.process(conv1, capi.STAGE_CONV)
    full1 = x.copy()
    eng.process(full1, capi.STAGE_ALL)
    eng.set_streaming(True)
    conv2 = _segmented(eng, x, seg * block, capi.STAGE_CONV)
    assert eng.stream_position() == T
    assert np.array_equal(conv1, conv2), np.abs(conv1 - conv2).max()
    eng.reset()
    full2 = _segmented(eng, x, seg * block, capi.STAGE_ALL)
    # a third pass: half-way the stream moves to another handle
    eng.reset()
    half = (n_cb // 2 // seg) * seg * block      # on a segment boundary of the run above, so that the EQ's scan tiles coincide
    a = _segmented(eng, x[:, :half], seg * block, capi.STAGE_ALL)
    blob = eng.export_state()
    eng.close()
    eng2, _ = _engine(sr, ir_len, T, block)
    eng2.set_streaming(True)
    eng2.import_state(blob)
    b = _segmented(eng2, x[:, half:], seg * block, capi.STAGE_ALL)
    eng2.close()
    assert np.abs(full1 - full2).max() <= 1e-12
    assert np.array_equal(np.concatenate([a, b], axis=1), full2)
    for s in range(2):
        want = checker.chain_run((irs[2 * s], irs[2 * s + 1]), signals.to_eqband(signals.band_params(80 + s)), x[2 * s:2 * s + 2], sr, block,
                                 OFilterSpec(sample_rate=sr), makeup=1.1, known_block=known)
        assert np.abs(full2[2 * s:2 * s + 2] - want).max() <= TOL, s


@pytest.mark.parametrize("structure", [0, 1])
@pytest.mark.parametrize("seg", [1, 9])
def test_mid_side_bands_are_carried(checker, structure, seg):
    """EQProcessor::filterState[2] / [3]: the Mid and Side rows' band states, Serial (node path) and Parallel structure, three
    streams with different Mid/Side band sets, through conv -> EQ -> epilogue; state moved to another handle half-way."""
    sr, block, ir_len, n_cb, n = 48000.0, 512, 20000, 60, 3
    T = n_cb * block
    modes = [[3, 0, 4, 1, 2] * 4, [0] * 20, [0, 4, 3, 0] * 5]
    x = np.stack([signals.noise(T, 500 + i, 0.3) for i in range(2 * n)])
    irs = [signals.synth_ir(ir_len, 600 + i) for i in range(2 * n)]

    def make():
        eng = ConvoPeqEngine(n, 2, sr, block, T, conv_boundary=capi.CONV_OUTER)
        for s in range(n):
            for ch in range(2):
                eng.set_impulse(s, ch, irs[2 * s + ch], 1.0, None)
            eng.set_eq(s, signals.to_band(signals.band_params(40 + s, modes=modes[s])), 0.2, 0.0, structure, s == 2)
        eng.set_epilogue(1.1, 0)
        return eng

    eng = make()
    one = x.copy()
    eng.process(one, capi.STAGE_ALL)
    eng.set_streaming(True)
    two = _segmented(eng, x, seg * block, capi.STAGE_ALL)
    eng.reset()
    half = (n_cb // 2 // seg) * seg * block
    a = _segmented(eng, x[:, :half], seg * block, capi.STAGE_ALL)
    blob = eng.export_state()
    eng.close()
    eng2 = make()
    eng2.set_streaming(True)
    eng2.import_state(blob)
    b = _segmented(eng2, x[:, half:], seg * block, capi.STAGE_ALL)
    eng2.close()
    assert np.abs(one - two).max() <= 1e-12
    assert np.array_equal(np.concatenate([a, b], axis=1), two)
    for s in range(n):
        want = checker.chain_run((irs[2 * s], irs[2 * s + 1]), signals.to_eqband(signals.band_params(40 + s, modes=modes[s])),
                                 x[2 * s:2 * s + 2], sr, block, None, makeup=1.1, structure=structure, agc=(s == 2))
        assert np.abs(two[2 * s:2 * s + 2] - want).max() <= TOL, s


@pytest.mark.parametrize("mix,bypass,delay", [(0.6, False, 700), (0.35, False, 5000), (1.0, True, 300), (0.0, False, 1234)])
def test_dry_path_delay_ring_is_carried(mix, bypass, delay):
    """ConvolverProcessor::process with mix < 1 (or bypassed): the latency-compensated dry path reads `delay` samples back, across
    call boundaries -- calls shorter and longer than the delay; state moved to another handle half-way."""
    sr, block, ir_len, n_cb, seg = 48000.0, 512, 20000, 48, 3
    T = n_cb * block
    x = np.stack([signals.noise(T, 800 + i, 0.3) for i in range(4)])
    irs = [signals.synth_ir(ir_len, 810 + i) for i in range(4)]

    def make():
        eng = ConvoPeqEngine(2, 2, sr, block, T, conv_boundary=capi.CONV_OUTER)
        for s in range(2):
            for ch in range(2):
                eng.set_impulse(s, ch, irs[2 * s + ch], 1.0, None)
            eng.set_eq(s, signals.to_band(signals.band_params(820 + s)))
        eng.set_epilogue(1.0, 0)
        eng.set_mix(mix, delay)
        eng.set_convolver_bypass(bypass)
        return eng

    eng = make()
    conv1 = x.copy()
    eng.process(conv1, capi.STAGE_CONV)
    one = x.copy()
    eng.process(one, capi.STAGE_ALL)
    eng.set_streaming(True)
    conv2 = _segmented(eng, x, seg * block, capi.STAGE_CONV)
    assert np.array_equal(conv1, conv2)
    eng.reset()
    two = _segmented(eng, x, seg * block, capi.STAGE_ALL)
    eng.reset()
    half = (n_cb // 2 // seg) * seg * block
    a = _segmented(eng, x[:, :half], seg * block, capi.STAGE_ALL)
    blob = eng.export_state()
    eng.close()
    eng2 = make()
    eng2.set_streaming(True)
    eng2.import_state(blob)
    b = _segmented(eng2, x[:, half:], seg * block, capi.STAGE_ALL)
    eng2.close()
    assert np.abs(one - two).max() <= 1e-12
    assert np.array_equal(np.concatenate([a, b], axis=1), two)
    if bypass or mix <= 0.001:      # the convolver stage is the delayed input alone
        assert np.array_equal(conv1[:, delay:], x[:, :T - delay]) and not conv1[:, :delay].any()


@pytest.mark.parametrize("block,seg,head", [(1024, 1, False), (2048, 3, False), (4096, 2, False), (512, 1, True), (512, 5, True)])
def test_larger_blocks_and_the_direct_form_head(checker, block, seg, head):
    """Blocks 1024 / 2048 / 4096 with the default FilterSpec: plans whose tail reader starves (skipped callbacks), P up to 32768
    (the four-step transforms); and the experimental direct-form head, whose FIR reaches 31 samples back across the call boundary."""
    sr, ir_len, n_cb = 48000.0, 131072, 72
    T = n_cb * block
    eng = ConvoPeqEngine(1, 2, sr, block, T)
    if head:
        eng.set_direct_head(True)
    spec = capi.default_filter_spec()
    irs = [signals.synth_ir(ir_len, 170 + ch) for ch in range(2)]
    for ch in range(2):
        eng.set_impulse(0, ch, irs[ch], 1.0, spec)
    x = np.stack([signals.noise(T, 180), signals.noise(T, 181)])
    one = x.copy()
    eng.process(one, capi.STAGE_CONV)
    eng.set_streaming(True)
    two = _segmented(eng, x, seg * block, capi.STAGE_CONV)
    eng.close()
    assert np.array_equal(one, two), np.abs(one - two).max()
    for ch in range(2):
        want, _ = checker.nuc_run(irs[ch], x[ch], block, spec=OFilterSpec(), direct_head=head)
        assert np.abs(two[ch] - want).max() <= TOL and np.abs(want).max() > 1e-3


@pytest.mark.parametrize("seg", [1, 3, 16])
def test_total_gain_ramp_across_calls(checker, seg):
    """EQProcessor::setTotalGain mid-stream: the 50 ms LinearRamp (4.7 callbacks) starts in one call and goes on in the next ones;
    at_callback counts from the call that follows the scheduling."""
    sr, block, n_cb, at = 48000.0, 512, 48, 20
    T = n_cb * block
    params = signals.band_params(seed=7)
    xl, xr = signals.log_sweep(T, sr)
    x = np.stack([xl, xr])
    eng = ConvoPeqEngine(1, 2, sr, block, T)
    eng.set_eq(0, signals.to_band(params), 0.2, 0.0)
    eng.schedule_total_gain(0, at, -6.0)
    one = x.copy()
    eng.process(one, capi.STAGE_EQ)
    eng.set_eq(0, signals.to_band(params), 0.2, 0.0)       # clears the schedule
    eng.set_streaming(True)
    two = np.empty_like(x)
    for c0 in range(0, n_cb, seg):
        c1 = min(n_cb, c0 + seg)
        if c0 <= at < c1:
            eng.schedule_total_gain(0, at - c0, -6.0)
        part = np.ascontiguousarray(x[:, c0 * block:c1 * block])
        eng.process(part, capi.STAGE_EQ)
        two[:, c0 * block:c1 * block] = part
    eng.close()
    wl, wr, _ = checker.eq_run(signals.to_eqband(params), xl, xr, sr, block, gain_change_db=-6.0, gain_change_at=at * block)
    assert np.abs(one - two).max() <= 1e-12
    assert np.abs(two[0] - wl).max() <= TOL and np.abs(two[1] - wr).max() <= TOL
    assert abs(two[0, -1] / one[0, -1] - 1.0) < 1e-9 and np.abs(two[:, (at + 6) * block:]).max() < 0.9 * np.abs(x).max() * 2


def test_tile_aligned_segments_are_bit_identical_for_the_whole_chain():
    """Segments of 16 callbacks = 8192 samples = one EQ scan tile: the tile grid of the segmented run coincides with the
    one-shot run's, so every stage gives the same bits."""
    sr, block, T = 48000.0, 512, 8192 * 6
    eng, _ = _engine(sr, 65536, T, block)
    x = np.stack([signals.noise(T, 190 + i) for i in range(4)])
    one = x.copy()
    eng.process(one, capi.STAGE_ALL)
    eng.set_streaming(True)
    two = _segmented(eng, x, 8192, capi.STAGE_ALL)
    eng.close()
    assert np.array_equal(one, two)


def test_three_layer_plan_block_64(checker):
    """Block 64: 64 / 512 / 4096 partitions, three layers, tails read through two delay lines."""
    sr, block, ir_len, T = 48000.0, 64, 65536, 64 * 600
    eng = ConvoPeqEngine(1, 2, sr, block, T)
    irs = [signals.synth_ir(ir_len, 3 + ch) for ch in range(2)]
    for ch in range(2):
        eng.set_impulse(0, ch, irs[ch])
    assert eng.layout().num_layers == 3
    x = np.stack([signals.noise(T, 5), signals.noise(T, 6)])
    one = x.copy()
    eng.process(one, capi.STAGE_CONV)
    eng.set_streaming(True)
    for seg in (3, 50, 129):
        eng.reset()
        two = _segmented(eng, x, seg * block, capi.STAGE_CONV)
        assert np.array_equal(one, two), seg
    eng.close()
    for ch in range(2):
        want, _ = checker.nuc_run(irs[ch], x[ch], block)
        assert np.abs(one[ch] - want).max() <= TOL


def test_all_stage_states_carry(checker):
    """AGC envelopes, OutputFilter / DC-blocker states and the limiter envelope across segment boundaries (segments of 5
    callbacks), against the one-shot call and the reference's output stages."""
    sr, block, T = 48000.0, 512, 512 * 120
    x = np.stack([signals.noise(T, 31, 0.35), signals.noise(T, 32, 0.35)])
    params = signals.band_params(seed=21)

    def make():
        eng = ConvoPeqEngine(1, 2, sr, block, T)
        eng.set_eq(0, signals.to_band(params), 0.2, 0.0, agc=True)
        eng.set_output_filter(True, conv_is_last=False)
        eng.set_output_stage(3.0, True)
        eng.set_peak_limiter(100.0)
        eng.set_epilogue(1.6, 0)
        return eng

    stages = capi.STAGE_EQ | capi.STAGE_OUTPUT_FILTER | capi.STAGE_EPILOGUE
    eng = make()
    one = x.copy()
    eng.process(one, stages)
    agc1 = eng.agc_state(0)
    eng.set_streaming(True)
    two = _segmented(eng, x, 5 * block, stages)
    agc2 = eng.agc_state(0)
    eng.close()
    assert np.abs(one - two).max() <= 1e-12
    assert np.abs(agc1 - agc2).max() <= 1e-12
    wl, wr, _ = checker.eq_run(signals.to_eqband(params), x[0], x[1], sr, block, agc=True)
    want = checker.output_run(np.stack([wl, wr]), sr, block, use_filter=True, conv_is_last=False, makeup=1.6, dc_cutoff=3.0, headroom=True,
                              clamp=True, limiter_ms=100.0)
    assert np.abs(two - want).max() <= TOL


def test_dither_history_carries_bit_for_bit(checker):
    """The shaper's 12-tap error history across calls: chaotic recurrence, so segmenting must not change a single bit."""
    sr, block, T, bits = 48000.0, 512, 512 * 40, 24
    x = np.stack([signals.noise(T, 41, 0.3), signals.noise(T, 42, 0.3)])
    u = np.random.default_rng(43).random((2, 2 * T))
    eng = ConvoPeqEngine(1, 2, sr, block, T)
    eng.set_streaming(True)
    y = np.empty_like(x)
    for t0 in range(0, T, 3 * block):
        t1 = min(T, t0 + 3 * block)
        eng.set_epilogue(0.9, bits, np.ascontiguousarray(u[:, 2 * t0:2 * t1]))
        part = np.ascontiguousarray(x[:, t0:t1])
        eng.process(part, capi.STAGE_EPILOGUE)
        y[:, t0:t1] = part
    eng.close()
    want, _ = checker.dither_run(x * 0.9, u, sr, bits, block)
    assert np.array_equal(y, want)


def test_dither_generator_state_carries(checker):
    """cpq_set_dither_seed in streaming mode: the xorshift64* state of every channel continues across calls."""
    sr, block, T, bits = 48000.0, 512, 512 * 30, 24
    x = np.stack([signals.noise(T, 45, 0.3), signals.noise(T, 46, 0.3)])
    eng = ConvoPeqEngine(1, 2, sr, block, T)
    eng.set_epilogue(1.0, bits)
    eng.set_dither_seed([12345])
    eng.set_streaming(True)
    y = _segmented(eng, x, 7 * block, capi.STAGE_EPILOGUE)
    eng.close()
    want, _ = checker.dither_run_seeded(x, 12345, sr, bits, block)
    assert np.array_equal(y, want)


def test_export_import_state_moves_a_stream_to_another_handle():
    sr, block, T = 48000.0, 512, 512 * 96
    x = np.stack([signals.noise(T, 51 + i) for i in range(4)])
    a, _ = _engine(sr, 65536, T, block)
    one = x.copy()
    a.process(one, capi.STAGE_ALL)
    a.set_streaming(True)
    half = 512 * 40
    first = np.ascontiguousarray(x[:, :half])
    a.process(first, capi.STAGE_ALL)
    blob = a.export_state()
    a.close()
    b, _ = _engine(sr, 65536, T, block)
    b.set_streaming(True)
    b.import_state(blob)
    assert b.stream_position() == half
    second = np.ascontiguousarray(x[:, half:])
    b.process(second, capi.STAGE_ALL)
    # a handle with another plan refuses the blob
    c, _ = _engine(sr, 30000, T, block)
    c.set_streaming(True)
    with pytest.raises(capi.CpqError):
        c.import_state(blob)
    c.close()
    b.close()
    assert np.abs(np.concatenate([first, second], axis=1) - one).max() <= 1e-12


def test_reset_starts_a_new_stream():
    sr, block, T = 48000.0, 512, 512 * 32
    eng, _ = _engine(sr, 65536, T, block, n_streams=1)
    x = np.stack([signals.noise(T, 61), signals.noise(T, 62)])
    eng.set_streaming(True)
    a = x.copy()
    eng.process(a, capi.STAGE_ALL)
    b = x.copy()
    eng.process(b, capi.STAGE_ALL)          # continues: the IR tail of the first pass is still ringing
    assert np.abs(a - b).max() > 1e-6
    eng.reset()
    c = x.copy()
    eng.process(c, capi.STAGE_ALL)
    eng.close()
    assert np.array_equal(a, c)
