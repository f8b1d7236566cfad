"""GPU parity at BASELINE.json's full sizes (configs[0..4]) through the C ABI.

Where the checker (the compiled reference when oracle/_ref exists, else the C restatement) finishes in seconds the whole
output is compared (cfg1, cfg2, cfg3, cfg5); the 1,024-stream batch (cfg4) is checked on sampled streams against the
checker plus a property over all of them: streams that were given identical IRs, EQs and inputs must produce
bit-identical outputs wherever they sit in the batch (different sequence chunks, CTAs and chain records).
Tolerance: max abs error <= 1e-10 of full scale (BASELINE.json north_star)."""
import numpy as np
import pytest

from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
from oracle.bindings import FilterSpec as OFilterSpec
from tests import signals

pytestmark = pytest.mark.gpu
TOL = 1e-10
BLOCK = 512


def _whole(n, block=BLOCK):
    return (n + block - 1) // block * block


def test_cfg1_stereo_65536_taps_10s_noise(checker):
    """configs[0]: stereo 48 kHz, 65,536-tap IR, block 512, 10 s white noise.  The reference cannot express a uniform
    partition beyond 16,384 taps: its own plan is L0 12x512 + L1 15x4096 (D1 7168, g1 1.4375), which is what runs."""
    sr, T = 48000.0, _whole(480000)
    irs = [signals.synth_ir(65536, 2 + ch) for ch in range(2)]
    x = np.stack([signals.noise(T, 1 + 10 * ch) for ch in range(2)])
    eng = ConvoPeqEngine(1, 2, sr, BLOCK, T, conv_boundary=capi.CONV_INNER)
    for ch in range(2):
        eng.set_impulse(0, ch, irs[ch], 1.0, None)
    lay = eng.layout()
    assert lay.num_layers == 2 and lay.layers[0].num_parts_ir == 12 and lay.layers[1].num_parts_ir == 15
    y = x.copy()
    eng.process(y, capi.STAGE_CONV)
    eng.close()
    for ch in range(2):
        want, _ = checker.nuc_run(irs[ch], x[ch], BLOCK)
        assert np.abs(y[ch] - want).max() <= TOL, ch


def test_cfg1b_uniform_extension_10s_noise():
    """configs[0] as worded: 65,536 taps in 128 uniform partitions of 512 (cfg.uniform_partitions, our extension): 10 s of
    noise against linear convolution."""
    from scipy.signal import fftconvolve
    sr, T = 48000.0, _whole(480000)
    irs = [signals.synth_ir(65536, 2 + ch) for ch in range(2)]
    x = np.stack([signals.noise(T, 1 + 10 * ch) for ch in range(2)])
    eng = ConvoPeqEngine(1, 2, sr, BLOCK, T, uniform_partitions=True)
    for ch in range(2):
        eng.set_impulse(0, ch, irs[ch], 1.0, None)
    lay = eng.layout()
    assert lay.num_layers == 1 and lay.layers[0].num_parts_ir == 128
    y = x.copy()
    eng.process(y, capi.STAGE_CONV)
    eng.close()
    for ch in range(2):
        lin = fftconvolve(x[ch], irs[ch])[:T]
        assert np.abs(y[ch] - lin).max() <= 1e-12 * max(1.0, np.abs(lin).max()), ch


@pytest.mark.parametrize("sat,stress", [(0.2, False), (0.0, False), (0.2, True), (0.0, True)])
def test_cfg2_eq_only_60s_sweep(checker, sat, stress):
    """configs[1]: stereo 48 kHz, 20-band peaking/shelf cascade only, 60 s log sweep; the mild set (U(-6,6) dB, Q U(0.5,4)) and
    SURVEY 8d's stress set (Q = 20, +-24 dB alternating: pole radius 0.99998, a memory of 6e4 samples across 352 tiles)."""
    sr, T = 48000.0, _whole(2880000)
    params = signals.band_params(seed=7, stress=stress)
    xl, xr = signals.log_sweep(T, sr)
    eng = ConvoPeqEngine(1, 2, sr, BLOCK, T)
    eng.set_eq(0, signals.to_band(params), sat, 0.0)
    y = np.stack([xl, xr]).copy()
    eng.process(y, capi.STAGE_EQ)
    state = eng.eq_state(0)
    eng.close()
    wl, wr, wstate = checker.eq_run(signals.to_eqband(params), xl, xr, sr, BLOCK, saturation=sat)
    assert np.abs(y[0] - wl).max() <= TOL and np.abs(y[1] - wr).max() <= TOL
    assert np.abs(state - wstate).max() <= 1e-9


def test_cfg3_96k_262144_taps_conv_then_eq(checker):
    """configs[2]: stereo 96 kHz, 262,144-tap IR, non-uniform partitions + 20-band EQ, Conv->EQ order, 10 s noise.
    At block 512 the reference's plan is L0 23x512 + L1 62x4096."""
    sr, T = 96000.0, _whole(960000)
    irs = [signals.synth_ir(262144, 20 + ch) for ch in range(2)]
    x = np.stack([signals.noise(T, 30 + ch) for ch in range(2)])
    params = signals.band_params(seed=7)
    eng = ConvoPeqEngine(1, 2, sr, BLOCK, T, conv_boundary=capi.CONV_OUTER)
    spec = capi.default_filter_spec(sample_rate=sr)
    for ch in range(2):
        eng.set_impulse(0, ch, irs[ch], 1.0, spec)
    lay = eng.layout()
    assert lay.layers[0].num_parts_ir == 23 and lay.layers[1].num_parts_ir == 62
    eng.set_eq(0, signals.to_band(params))
    eng.set_epilogue(1.0, 0)
    y = x.copy()
    eng.process(y, capi.STAGE_ALL)
    eng.close()
    want = checker.chain_run((irs[0], irs[1]), signals.to_eqband(params), x, sr, BLOCK, OFilterSpec(sample_rate=sr))
    assert np.abs(y - want).max() <= TOL


def test_cfg4_batch_of_1024_stereo_streams(checker):
    """configs[3]: 1,024 independent stereo streams at 48 kHz, 131,072-tap IRs + EQ, 10 s each, device-resident.
    Streams s and s + 512 share IRs, EQ and input: their outputs must be bit-identical (all 512 pairs); eight sampled
    streams are compared with the checker's conv -> EQ -> makeup/headroom chain."""
    import torch
    sr, T, S, ir_len = 48000.0, _whole(480000), 1024, 131072
    half = S // 2
    dev = torch.device("cuda", 0)
    eng = ConvoPeqEngine(S, 2, sr, BLOCK, T, device=0, conv_boundary=capi.CONV_OUTER)
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    decay = torch.exp(-torch.arange(ir_len, device=dev, dtype=torch.float64) / (ir_len / 6.0)) / (ir_len ** 0.5)
    spec = capi.default_filter_spec()
    sampled = [0, 1, 58, 59, 60, 255, 300, 511]      # incl. both sides of the first sequence-chunk boundary (118 sequences)
    kept_ir = {}
    for s0 in range(0, 2 * half, 64):
        irs = (torch.randn(64, ir_len, device=dev, dtype=torch.float64, generator=g) * decay).cpu().numpy()
        for i in range(64):
            q = s0 + i
            for rep in (0, half):
                eng.set_impulse(q // 2 + rep, q % 2, irs[i], 1.0, spec)
            if q // 2 in sampled:
                kept_ir[q] = irs[i].copy()
    for s in range(half):
        b = signals.to_band(signals.band_params(1000 + s))
        eng.set_eq(s, b, 0.2, 0.0)
        eng.set_eq(s + half, b, 0.2, 0.0)
    eng.set_epilogue(1.0, 0)
    x = torch.empty(2 * S, T, device=dev, dtype=torch.float64)
    x[:2 * half] = torch.randn(2 * half, T, device=dev, dtype=torch.float64, generator=g) * 0.1
    x[2 * half:] = x[:2 * half]
    x_sampled = {s: x[2 * s:2 * s + 2].cpu().numpy() for s in sampled}
    # the same batch through the host entry point (cpq_process: chunked H2D / compute / D2H on three streams, the e2e
    # number's path) must give the same bits as the device-resident call
    host = torch.empty(2 * S, T, dtype=torch.float64).pin_memory()
    host.copy_(x)
    eng.process_device(x.data_ptr(), T, T, capi.STAGE_ALL)
    torch.cuda.synchronize()
    eng.process_host_ptrs(host.data_ptr(), T, T, capi.STAGE_ALL)
    assert torch.equal(host, x.cpu()), "cpq_process (host buffers) differs from cpq_process_device"
    del host
    assert torch.equal(x[:2 * half], x[2 * half:])
    assert bool(torch.isfinite(x).all())
    for s in sampled:
        want = checker.chain_run((kept_ir[2 * s], kept_ir[2 * s + 1]), signals.to_eqband(signals.band_params(1000 + s)),
                                 x_sampled[s], sr, BLOCK, OFilterSpec())
        got = x[2 * s:2 * s + 2].cpu().numpy()
        assert np.abs(got - want).max() <= TOL, s
    eng.close()


def test_cfg5_192k_8ch_2M_taps(checker):
    """configs[4]: one 192 kHz 8-channel stream with a 2,097,152-tap IR (reference hard maximum), 10 s: L0 32x512 +
    L1 64x4096 + L2 56x32768 (D2 60416, g2 1.1), all partitions on one GPU (the partition-range split over ranks is
    covered by test_partition_range_partials_sum_to_full and scripts/multi_gpu_check.py)."""
    sr, T, n_ch = 192000.0, _whole(1920000), 8
    ir = [signals.synth_ir(2097152, 40 + ch % 2) for ch in range(2)]
    x = np.stack([signals.noise(T, 50 + ch) for ch in range(n_ch)])
    eng = ConvoPeqEngine(n_ch // 2, 2, sr, BLOCK, T, conv_boundary=capi.CONV_INNER, shared_ir=True)
    spec = capi.default_filter_spec(sample_rate=sr)
    for ch in range(2):
        eng.set_impulse(-1, ch, ir[ch], 1.0, spec)
    lay = eng.layout()
    assert lay.num_layers == 3 and lay.layers[2].part_size == 32768 and lay.layers[2].num_parts_ir == 56
    y = x.copy()
    eng.process(y, capi.STAGE_CONV)
    eng.close()
    ospec = OFilterSpec(sample_rate=sr)
    for ch in (0, 1, 5, 6):
        want, _ = checker.nuc_run(ir[ch % 2], x[ch], BLOCK, spec=ospec)
        assert np.abs(y[ch] - want).max() <= TOL, ch
