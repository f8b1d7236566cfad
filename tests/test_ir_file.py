"""IR files the way the application loads them: decode, trailing-silence trim, mixed phase, the loader pipeline.

The decode restates JUCE's WavAudioFormat reader (external to the reference tree) as LoaderThread::doLoadStep uses it
(convolver/ConvolverProcessor.LoaderThread.cpp:439-486): everything goes through 32-bit float, then the input transform."""
import io
import os
import struct

import numpy as np
import pytest

from convopeq_b200 import capi, engine
from tests import signals

REF_SAMPLES = "/root/reference/sampledata"


def _wav(fmt_tag, bits, channels, rate, payload, extensible=False):
    align = channels * bits // 8
    if extensible:
        guid = struct.pack("<H", fmt_tag) + bytes.fromhex("000000001000800000aa00389b71")
        fmt = struct.pack("<HHIIHHHHI", 0xFFFE, channels, rate, rate * align, align, bits, 22, bits, 3) + guid
    else:
        fmt = struct.pack("<HHIIHH", fmt_tag, channels, rate, rate * align, align, bits)
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt
    chunks += b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\0"            # an odd-sized chunk before the data (pad byte)
    chunks += b"data" + struct.pack("<I", len(payload)) + payload
    return b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks


def _juce_float(ints32):
    """int32 (left-justified) -> float the way AudioFormatReader::read does: (float) i * (1.0f / 0x7fffffff)."""
    return (ints32.astype(np.float32) * np.float32(1.0 / np.float32(0x7fffffff))).astype(np.float64)


def _transform(x):
    y = x.copy()
    n = y.shape[-1]
    tail = np.arange(n) >= n // 4 * 4
    y[np.isnan(y) | (np.abs(y) < 1e-20) | (np.isinf(y) & tail)] = 0.0
    return np.clip(y, -1.0, 1.0)


@pytest.mark.parametrize("kind", ["pcm8", "pcm16", "pcm24", "pcm32", "f32", "f64", "pcm24x", "f32x"])
def test_decode_matches_the_juce_rule(kind):
    g = np.random.default_rng(5)
    ch, frames, rate = 2, 1001, 44100
    ext = kind.endswith("x")
    k = kind.rstrip("x")
    if k == "pcm8":
        raw = g.integers(0, 256, (frames, ch), dtype=np.uint8)
        data, want = _wav(1, 8, ch, rate, raw.tobytes(), ext), _juce_float((raw.astype(np.int64) - 128 << 24).astype(np.int32))
    elif k == "pcm16":
        raw = g.integers(-32768, 32768, (frames, ch), dtype=np.int16)
        data, want = _wav(1, 16, ch, rate, raw.tobytes(), ext), _juce_float(raw.astype(np.int32) << 16)
    elif k == "pcm24":
        raw = g.integers(-(1 << 23), 1 << 23, (frames, ch), dtype=np.int32)
        b = (raw.astype(np.uint32) & 0xFFFFFF).astype("<u4").view(np.uint8).reshape(frames, ch, 4)[:, :, :3]
        data, want = _wav(1, 24, ch, rate, b.tobytes(), ext), _juce_float(raw << 8)
    elif k == "pcm32":
        raw = g.integers(-(1 << 31), 1 << 31, (frames, ch), dtype=np.int64).astype(np.int32)
        data, want = _wav(1, 32, ch, rate, raw.tobytes(), ext), _juce_float(raw)
    elif k == "f32":
        raw = (g.standard_normal((frames, ch)) * 0.5).astype(np.float32)
        raw[3, 0], raw[7, 1], raw[frames - 1, 0], raw[11, 1] = np.nan, 1e-30, np.inf, -3.0
        data, want = _wav(3, 32, ch, rate, raw.tobytes(), ext), raw.astype(np.float64)
    else:
        raw = g.standard_normal((frames, ch)) * 0.3
        data, want = _wav(3, 64, ch, rate, raw.tobytes(), ext), raw.astype(np.float32).astype(np.float64)
    got, sr, bits, is_float = engine.ir_decode_wav(data)
    assert sr == rate and got.shape == (ch, frames) and is_float == k.startswith("f")
    assert np.array_equal(got, _transform(want.T))


def test_decode_rejects_what_it_does_not_read():
    with pytest.raises(capi.CpqError):
        engine.ir_decode_wav(b"RIFF\0\0\0\0AIFFxxxx")
    with pytest.raises(capi.CpqError):
        engine.ir_decode_wav(_wav(2, 4, 1, 48000, b"\0" * 64))      # ADPCM


@pytest.mark.skipif(not os.path.isdir(REF_SAMPLES), reason="reference tree not mounted")
def test_decode_of_the_reference_sample_files_matches_scipy():
    from scipy.io import wavfile
    for name in ("impulse_room_correction_hpf_lpf.wav", "synthetic_long_ir_20s.wav"):
        path = os.path.join(REF_SAMPLES, name)
        rate, ref = wavfile.read(path)
        ref = ref.reshape(len(ref), -1)
        want = ref.astype(np.float64) if ref.dtype.kind == "f" else ref.astype(np.float64) / 32768.0
        got, sr, _, _ = engine.ir_decode_wav(open(path, "rb").read())
        assert sr == rate and got.shape == want.T.shape
        assert np.array_equal(got, _transform(want.T))


def test_trim_silence():
    a = np.zeros(1000)
    b = np.zeros(1000)
    a[10], b[500], b[501] = 1.0, 2e-15, 1e-15
    assert engine.ir_trim_silence(a) == 11
    assert engine.ir_trim_silence(a, b) == 501
    assert engine.ir_trim_silence(np.zeros(64), np.zeros(64)) == 1


def _mixed_numpy(lin, mnp, sr, lo, hi):
    """convertToMixedPhaseFallback (ConvolverProcessor.MixedPhase.cpp:721-865) in numpy."""
    L = len(lin)
    n = 1 << (L - 1).bit_length()
    zl, zm = np.fft.fft(lin, n), np.fft.fft(mnp, n)
    half = n // 2
    k = np.arange(half + 1)
    f = k * sr / n
    wmin = np.where(f >= hi, 0.0, np.where(f > lo, 0.5 * (1 + np.cos(np.pi * (f - lo) / (hi - lo))), 1.0))
    peak = int(np.argmax(np.abs(lin)))
    philin = -(2 * np.pi * k / n) * peak
    d = ((1 - wmin) * philin + wmin * np.angle(zm[:half + 1])) - philin
    corr = 0.0
    for i in range(1, len(d)):
        delta = d[i] - d[i - 1]
        if delta > np.pi:
            corr -= 2 * np.pi
        elif delta < -np.pi:
            corr += 2 * np.pi
        d[i] += corr
    full = np.concatenate([d, -d[1:half][::-1]])
    out = np.fft.ifft(zl * np.exp(1j * full)).real[:L]
    out[np.abs(out) < 1e-18] = 0.0
    return out


@pytest.mark.parametrize("L", [4096, 5000])
def test_mixed_phase_fallback(L):
    ir = np.roll(signals.synth_ir(L, 3), 700)      # a linear-phase-like IR: peak away from zero
    mnp = engine.ir_min_phase(ir)
    got = engine.ir_mixed_phase(ir, mnp, 48000.0)
    want = _mixed_numpy(ir, mnp, 48000.0, 200.0, 1000.0)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    # above the transition the phase is the linear IR's: the high band of the two spectra agrees, the low band does not
    n = 1 << (L - 1).bit_length()
    hb = slice(int(3000 / 48000 * n), n // 2)
    # (the kept L of n samples of the circular result: a small truncation error remains for L < n)
    tol = 1e-9 if L == n else 5e-2
    assert np.abs(np.fft.fft(got, n)[hb] - np.fft.fft(ir, n)[hb]).max() <= tol * np.abs(np.fft.fft(ir, n)).max()
    assert engine.ir_mixed_phase(ir, mnp, 48000.0, 1000.0, 1000.0) is None


@pytest.mark.gpu
@pytest.mark.parametrize("phase_mode,mono_file", [(0, False), (1, False), (2, True)])
def test_load_impulse_wav_pipeline(checker, phase_mode, mono_file):
    """decode -> trim -> prepare -> phase -> scale -> SetImpulse, against the same steps taken one by one, and the convolution
    of a signal against the checker with that IR."""
    from oracle.bindings import FilterSpec as OFilterSpec
    sr, block, T = 48000.0, 512, 16384
    g = np.random.default_rng(9)
    frames = 30000
    ir = (g.standard_normal((frames, 1 if mono_file else 2)) * np.exp(-np.arange(frames) / 4000.0)[:, None] * 0.2).astype(np.float32)
    ir[200] += 0.7
    ir[25000:] = 0.0                                   # trailing silence the loader drops
    data = _wav(3, 32, ir.shape[1], int(sr), ir.tobytes())
    eng = engine.ConvoPeqEngine(1, 2, sr, block, T)
    info = eng.load_impulse_wav(0, data, phase_mode, 0.25, capi.default_filter_spec())
    assert (info.file_channels, info.file_frames, info.trimmed_frames) == (ir.shape[1], frames, 25000)
    assert info.target_length == 12000 and info.phase_applied == phase_mode
    # the same steps one by one
    dec, _, _, _ = engine.ir_decode_wav(data)
    chans = [engine.ir_prepare(dec[c][:info.trimmed_frames], sr, 0.25) for c in range(dec.shape[0])]
    if phase_mode >= 1:
        mp = [engine.ir_min_phase(c) for c in chans]
        chans = mp if phase_mode == 1 else [engine.ir_mixed_phase(c, m, sr) for c, m in zip(chans, mp)]
    scale, has, _ = engine.ir_scale_factor(chans[0], chans[1] if len(chans) > 1 else None)
    assert has and abs(info.scale_factor - scale) <= 1e-15 * scale
    assert info.peak_latency == engine.ir_peak_latency(chans[0], chans[1] if len(chans) > 1 else None)
    x = np.stack([signals.noise(T, 1), signals.noise(T, 2)])
    y = x.copy()
    eng.process(y, capi.STAGE_CONV)
    eng.close()
    for ch in range(2):
        want, _ = checker.nuc_run(chans[min(ch, len(chans) - 1)], x[ch], block, scale=scale, spec=OFilterSpec())
        assert np.abs(y[ch] - want).max() <= 1e-10
