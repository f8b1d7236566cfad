"""Host-side mirror of the reference's processing interface over the C ABI (include/cpq.h).

Names follow the reference: prepare with IR, sample rate and block size (`prepare_to_play`,
`set_impulse` ~ StereoConvolver::init / MKLNonUniformConvolver::SetImpulse), set per-band parameters
(`set_band` ~ EQProcessor::setBand*/createCoeffCache), process a block in place (`process`).
PyTorch is only used by callers for device memory; nothing here imports it.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import capi

_dp = C.POINTER(C.c_double)

# EQProcessor::DEFAULT_FREQS (eqprocessor/EQProcessor.h:158-163): the processor's own band centres
EQPROCESSOR_DEFAULT_FREQS = (25.0, 40.0, 63.0, 100.0, 160.0, 250.0, 400.0, 630.0, 1000.0, 1600.0,
                             2500.0, 4000.0, 6300.0, 10000.0, 11000.0, 12500.0, 14000.0, 16500.0, 18000.0, 19500.0)

LOW_SHELF, PEAKING, HIGH_SHELF, LOW_PASS, HIGH_PASS = range(5)
STEREO, LEFT, RIGHT = range(3)


@dataclass
class Band:
    """convo::EQBandParams (core/EQParameters.h:11-20); float parameters, like the reference."""
    frequency: float = 1000.0
    gain: float = 0.0
    q: float = 0.707
    enabled: bool = True
    type: int = PEAKING
    channel_mode: int = STEREO


def design_band(band: Band, sample_rate: float) -> capi.SvfCoeffs:
    """EQProcessor::calcSVFCoeffs (EQProcessor.Coefficients.cpp:101-130)."""
    out = capi.SvfCoeffs()
    st = capi.load().cpq_design_band(int(band.type), C.c_float(band.frequency), C.c_float(band.gain),
                                     C.c_float(band.q), float(sample_rate), C.byref(out))
    if st != capi.OK:
        raise capi.CpqError(st, "design_band")
    return out


class ConvoPeqEngine:
    """A batch of independent (stereo or mono) streams on one B200."""

    def __init__(self, n_streams: int = 1, n_channels: int = 2, sample_rate: float = 48000.0, block_size: int = 512,
                 max_samples: int = 480000 // 512 * 512, device: int = 0, conv_boundary: int = capi.CONV_INNER,
                 shared_ir: bool = False, shared_eq: bool = False, workspace_bytes: int = 0, uniform_partitions: bool = False):
        self.lib = capi.load()
        cfg = capi.Config()
        self.lib.cpq_config_default(C.byref(cfg))
        cfg.device = device
        cfg.n_streams = n_streams
        cfg.n_channels = n_channels
        cfg.block_size = block_size
        cfg.sample_rate = sample_rate
        cfg.max_samples = max_samples
        cfg.conv_boundary = conv_boundary
        cfg.shared_ir = int(shared_ir)
        cfg.shared_eq = int(shared_eq)
        cfg.workspace_bytes = workspace_bytes
        cfg.uniform_partitions = int(uniform_partitions)
        self.cfg = cfg
        self.h = C.c_void_p()
        st = self.lib.cpq_create(C.byref(cfg), C.byref(self.h))
        if st != capi.OK:
            raise capi.CpqError(st, self.lib.cpq_last_error(None).decode())
        self.n_seq = n_streams * n_channels
        self.sample_rate = sample_rate

    # ---- lifecycle ----
    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.cpq_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int):
        if st != capi.OK:
            raise capi.CpqError(st, self.lib.cpq_last_error(self.h).decode())

    def reset(self):
        self._check(self.lib.cpq_reset(self.h))

    # ---- prepare ----
    def set_impulse(self, stream: int, channel: int, ir: np.ndarray, scale: float = 1.0,
                    spec: Optional[capi.FilterSpec] = None):
        ir = np.ascontiguousarray(ir, dtype=np.float64)
        self._check(self.lib.cpq_set_impulse(self.h, stream, channel, ir.ctypes.data_as(_dp), ir.size, scale,
                                             C.byref(spec) if spec is not None else None))

    def set_eq(self, stream: int, bands: Sequence[Band], saturation: float = 0.2, total_gain_db: float = 0.0,
               structure: int = 0, agc: bool = False):
        """createCoeffCache(params) + EQParameters (ProcessingCache.cpp:56-96). saturation/total gain are float in
        the reference and promoted to double before use.  structure: 0 Serial, 1 Parallel
        (EQParameters::filterStructure); agc: EQParameters::agcEnabled.  Band channel modes 3/4 are Mid/Side."""
        assert len(bands) == capi.NUM_BANDS
        co = (capi.SvfCoeffs * capi.NUM_BANDS)()
        act = (C.c_uint8 * capi.NUM_BANDS)()
        mode = (C.c_int32 * capi.NUM_BANDS)()
        for i, b in enumerate(bands):
            on = bool(b.enabled) and self.sample_rate > 0
            act[i] = 1 if on else 0
            mode[i] = int(b.channel_mode)
            co[i] = design_band(b, self.sample_rate) if on else capi.SvfCoeffs()
        # EQProcessor::setNonlinearSaturation clamps to [0, 1], setTotalGain to +-48 dB (EQProcessor.Parameters.cpp:206-208, 103-106)
        sat = float(np.float32(min(1.0, max(0.0, saturation))))
        gain = self.lib.cpq_db_to_gain(C.c_float(min(48.0, max(-48.0, total_gain_db))))
        self._check(self.lib.cpq_set_eq(self.h, stream, co, act, mode, sat, gain))
        node = (C.c_uint8 * capi.NUM_BANDS)(*[self.lib.cpq_band_node_active(int(b.type), C.c_float(b.gain), int(bool(b.enabled)),
                                                                             self.sample_rate) for b in bands])
        self._check(self.lib.cpq_set_eq_mode(self.h, stream, int(structure), int(agc), node))

    def agc_state(self, stream: int) -> np.ndarray:
        """(envIn, envOut, gain) of the stream's AGC after the last process call."""
        out = np.zeros(3)
        self._check(self.lib.cpq_get_agc_state(self.h, stream, out.ctypes.data_as(_dp)))
        return out

    def set_eq_raw(self, stream: int, coeffs: np.ndarray, active: Sequence[int], modes: Sequence[int],
                   saturation: float, total_gain_lin: float):
        co = (capi.SvfCoeffs * capi.NUM_BANDS)()
        for i in range(capi.NUM_BANDS):
            co[i] = capi.SvfCoeffs(*[float(v) for v in coeffs[i]])
        act = (C.c_uint8 * capi.NUM_BANDS)(*[int(a) for a in active])
        mode = (C.c_int32 * capi.NUM_BANDS)(*[int(m) for m in modes])
        self._check(self.lib.cpq_set_eq(self.h, stream, co, act, mode, saturation, total_gain_lin))

    def schedule_total_gain(self, stream: int, at_callback: int, gain_db: float):
        self._check(self.lib.cpq_schedule_total_gain(self.h, stream, at_callback,
                                                     self.lib.cpq_db_to_gain(C.c_float(gain_db))))

    def set_epilogue(self, makeup_gain: float = 1.0, dither_bits: int = 0, uniforms: Optional[np.ndarray] = None):
        self._check(self.lib.cpq_set_epilogue(self.h, makeup_gain, dither_bits))
        if uniforms is not None:
            u = np.ascontiguousarray(uniforms, dtype=np.float64)
            assert u.shape[0] == self.n_seq and u.shape[1] % 2 == 0
            self._check(self.lib.cpq_set_dither_uniforms(self.h, u.ctypes.data_as(_dp), u.shape[1] // 2))

    def set_dither_seed(self, stream_seeds: Optional[Sequence[int]]):
        """Dither uniforms from the reference's fallback generator (PsychoacousticDither(seed) per stream); None = injected uniforms."""
        if stream_seeds is None:
            self._check(self.lib.cpq_set_dither_seed(self.h, None))
            return
        assert len(stream_seeds) == self.cfg.n_streams
        arr = (C.c_uint64 * len(stream_seeds))(*[int(s) & 0xFFFFFFFFFFFFFFFF for s in stream_seeds])
        self._check(self.lib.cpq_set_dither_seed(self.h, arr))

    def set_dither_uniforms_device(self, data_ptr: int, samples_per_channel: int):
        """Injected dither uniforms already on the device: [n_seq][2 * samples] doubles, borrowed."""
        self._check(self.lib.cpq_set_dither_uniforms_device(self.h, data_ptr, samples_per_channel))

    def set_partition_range(self, begin: int, end: int):
        self._check(self.lib.cpq_set_partition_range(self.h, begin, end))

    def set_partial_sources(self, device_ptrs: Sequence[int]):
        """Every rank's partial buffer (device addresses valid in this process, rank order, own buffer included); [] clears."""
        arr = (C.c_void_p * max(len(device_ptrs), 1))(*[C.c_void_p(int(p)) for p in device_ptrs])
        self._check(self.lib.cpq_set_partial_sources(self.h, len(device_ptrs), arr))

    def set_stream_window(self, first_stream: int = 0, n_streams: int = -1):
        self._check(self.lib.cpq_set_stream_window(self.h, first_stream, n_streams))

    def total_partitions(self) -> int:
        return self.lib.cpq_total_partitions(self.h)

    # ---- process ----
    def set_output_filter(self, enabled: bool = True, conv_is_last: bool = False, hc_mode: int = 1, lc_mode: int = 0, lp_mode: int = 1):
        """OutputFilter::process(block, convIsLast, hcMode, lcMode, lpMode) (OutputFilter.h:108-131); runs with STAGE_OUTPUT_FILTER."""
        self._check(self.lib.cpq_set_output_filter(self.h, int(enabled), int(conv_is_last), hc_mode, lc_mode, lp_mode))

    def set_direct_head(self, enable: bool = True):
        """enableDirectHead of SetImpulse (experimental direct-form head); call before set_impulse."""
        self._check(self.lib.cpq_set_direct_head(self.h, int(enable)))

    def set_mix(self, mix: float, dry_delay_samples: int):
        """ConvolverProcessor's dry/wet mix (float mixTarget) and the latency-compensation delay of its dry path."""
        self._check(self.lib.cpq_set_mix(self.h, C.c_float(mix), int(dry_delay_samples)))

    def set_input_gain(self, gain: float = 1.0):
        """Gain of the engine's input stage (STAGE_INPUT: gain, NaN / denormal scrub, clamp to [-1, 1])."""
        self._check(self.lib.cpq_set_input_gain(self.h, gain))

    def set_peak_limiter(self, release_ms: float = 100.0):
        """SimplePeakLimiter between the scrub and the hard clamp (the reference engine prepares it with 100 ms); 0 = off."""
        self._check(self.lib.cpq_set_peak_limiter(self.h, release_ms))

    def set_convolver_bypass(self, bypassed: bool = True):
        """ConvolverProcessor bypass: the convolver stage becomes the latency-compensating delay (set_mix's dry delay)."""
        self._check(self.lib.cpq_set_convolver_bypass(self.h, int(bypassed)))

    def set_conv_input_trim(self, gain: float):
        """convolverInputTrimGain of the EQThenConvolver order (DSPCoreDouble.cpp:438-445)."""
        self._check(self.lib.cpq_set_conv_input_trim(self.h, gain))

    def set_output_stage(self, dc_cutoff_hz: float = 3.0, hard_clamp: bool = True):
        """Output DC blocker (AudioEngine.h:643-651 uses 3 Hz) and the scrub + +-kOutputHeadroom clamp of processOutputDouble."""
        self._check(self.lib.cpq_set_output_stage(self.h, dc_cutoff_hz, int(hard_clamp)))

    def load_impulse_wav(self, stream: int, data: bytes, phase_mode: int = 0, target_seconds: float = 1.0, spec=None) -> capi.IrLoadInfo:
        """ConvolverProcessor::loadImpulseResponse for one stream from the bytes of a WAV file at the engine's rate: decode, trim,
        DC blocker / window / target length, phase mode (0 as is, 1 minimum, 2 mixed), scale factor, peak latency, SetImpulse."""
        info = capi.IrLoadInfo()
        self._check(self.lib.cpq_load_impulse_wav(self.h, stream, data, len(data), phase_mode, target_seconds,
                                                  C.byref(spec) if spec is not None else None, C.byref(info)))
        return info

    def process(self, x: np.ndarray, stages: int = capi.STAGE_ALL) -> np.ndarray:
        """In place on a host array [n_seq, T] (rows = stream*n_channels + ch). H2D/D2H inside."""
        assert x.dtype == np.float64 and x.ndim == 2 and x.shape[0] == self.n_seq and x.flags["C_CONTIGUOUS"]
        ptrs = (_dp * self.n_seq)(*[x[i].ctypes.data_as(_dp) for i in range(self.n_seq)])
        self._check(self.lib.cpq_process(self.h, ptrs, x.shape[1], stages))
        return x

    def process_f32(self, x: np.ndarray, stages: int = capi.STAGE_ALL) -> np.ndarray:
        """In place on a float32 host array [n_seq, T]: FP32 on the wire, FP64 arithmetic."""
        assert x.dtype == np.float32 and x.ndim == 2 and x.shape[0] == self.n_seq and x.flags["C_CONTIGUOUS"]
        fp = C.POINTER(C.c_float)
        ptrs = (fp * self.n_seq)(*[x[i].ctypes.data_as(fp) for i in range(self.n_seq)])
        self._check(self.lib.cpq_process_f32(self.h, ptrs, x.shape[1], stages))
        return x

    def process_f32_host_ptrs(self, base_ptr: int, row_stride_floats: int, T: int, stages: int = capi.STAGE_ALL):
        fp = C.POINTER(C.c_float)
        ptrs = (fp * self.n_seq)(*[C.cast(base_ptr + 4 * i * row_stride_floats, fp) for i in range(self.n_seq)])
        self._check(self.lib.cpq_process_f32(self.h, ptrs, T, stages))

    def process_host_ptrs(self, base_ptr: int, row_stride_doubles: int, T: int, stages: int = capi.STAGE_ALL):
        """Same as process() for a host buffer given by address (e.g. a pinned torch tensor)."""
        ptrs = (_dp * self.n_seq)(*[C.cast(base_ptr + 8 * i * row_stride_doubles, _dp) for i in range(self.n_seq)])
        self._check(self.lib.cpq_process(self.h, ptrs, T, stages))

    def process_device(self, data_ptr: int, stride: int, T: int, stages: int = capi.STAGE_ALL):
        """In place on device memory [n_seq][stride] doubles (e.g. torch tensor .data_ptr())."""
        self._check(self.lib.cpq_process_device(self.h, C.c_void_p(data_ptr), stride, T, stages))

    # ---- introspection ----
    def layout(self) -> capi.Layout:
        out = capi.Layout()
        self._check(self.lib.cpq_get_layout(self.h, C.byref(out)))
        return out

    def latency(self) -> int:
        return self.lib.cpq_latency(self.h)

    def timings(self) -> capi.Timings:
        t = capi.Timings()
        self._check(self.lib.cpq_get_timings(self.h, C.byref(t)))
        return t

    # ---- streaming continuation (cpq_set_streaming): Add/Get/EQ state carried between calls ----
    def set_streaming(self, enable: bool = True):
        self._check(self.lib.cpq_set_streaming(self.h, int(enable)))

    def stream_position(self) -> int:
        return int(self.lib.cpq_stream_position(self.h))

    def export_state(self) -> np.ndarray:
        n = int(self.lib.cpq_state_size(self.h))
        blob = np.empty(n, dtype=np.uint8)
        self._check(self.lib.cpq_export_state(self.h, blob.ctypes.data, n))
        return blob

    def import_state(self, blob: np.ndarray):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self._check(self.lib.cpq_import_state(self.h, blob.ctypes.data, blob.size))

    def eq_state(self, stream: int) -> np.ndarray:
        out = np.zeros((self.cfg.n_channels, capi.NUM_BANDS, 2))
        self._check(self.lib.cpq_get_eq_state(self.h, stream, out.ctypes.data_as(_dp)))
        return out

    def cuda_stream(self) -> int:
        return int(self.lib.cpq_cuda_stream(self.h) or 0)

    def kernel_launch_count(self) -> int:
        return int(self.lib.cpq_kernel_launch_count(self.h))


# convo::EQParameters defaults (core/EQParameters.h:32-37): the state a preset is loaded into
EQPARAMETERS_DEFAULT_FREQS = (20.0, 32.0, 50.0, 80.0, 125.0, 200.0, 315.0, 500.0, 800.0, 1250.0, 2000.0, 3150.0, 5000.0, 8000.0, 12500.0,
                              16000.0, 19000.0, 20000.0, 22000.0, 24000.0)


def ir_prepare(ir: np.ndarray, sample_rate: float, target_seconds: float) -> np.ndarray:
    """One channel of a loaded IR (at the device rate) through the reference's DC blocker / Tukey window / trim + fade."""
    lib = capi.load()
    ir = np.ascontiguousarray(ir, dtype=np.float64)
    n = lib.cpq_ir_target_length(sample_rate, target_seconds)
    out = np.zeros(n)
    got = lib.cpq_ir_prepare(ir.ctypes.data_as(_dp), ir.size, sample_rate, target_seconds, out.ctypes.data_as(_dp), n)
    if got != n:
        raise ValueError("cpq_ir_prepare")
    return out


def load_eq_preset(text: str, bands: Optional[Sequence[Band]] = None, total_gain_db: float = 0.0):
    """EQProcessor::loadFromTextFile on the contents of an EqualizerAPO / AutoEq preset; returns (bands, total_gain_db,
    ignored_filter_lines).  `bands` is the state before loading (default: convo::EQParameters{})."""
    arr = (capi.EqBandParams * capi.NUM_BANDS)()
    start = list(bands) if bands is not None else [Band(frequency=f) for f in EQPARAMETERS_DEFAULT_FREQS]
    for i, b in enumerate(start):
        arr[i] = capi.EqBandParams(b.frequency, b.gain, b.q, int(b.enabled), int(b.type), int(b.channel_mode))
    g = C.c_float(total_gain_db)
    ignored = capi.load().cpq_parse_eq_preset(text.encode("utf-8", "replace"), arr, C.byref(g))
    if ignored < 0:
        raise ValueError("cpq_parse_eq_preset")
    out = [Band(a.frequency, a.gain_db, a.q, bool(a.enabled), a.type, a.channel_mode) for a in arr]
    return out, float(g.value), ignored


def ir_scale_factor(ir_l: np.ndarray, ir_r: Optional[np.ndarray] = None, cur_l: Optional[np.ndarray] = None,
                    cur_r: Optional[np.ndarray] = None, cur_scale: float = 1.0):
    """IRConverter::computeScaleFactor (host-only) -> (scale_factor, has_scale_factor, additional_attenuation_db)."""
    arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (ir_l, ir_r, cur_l, cur_r)]
    ptr = lambda a: a.ctypes.data_as(_dp) if a is not None else None
    out = capi.IrScale()
    st = capi.load().cpq_ir_scale_factor(ptr(arrs[0]), ptr(arrs[1]), arrs[0].size, ptr(arrs[2]), ptr(arrs[3]),
                                         0 if arrs[2] is None else arrs[2].size, cur_scale, C.byref(out))
    if st != capi.OK:
        raise capi.CpqError(st, "cpq_ir_scale_factor")
    return float(out.scale_factor), bool(out.has_scale_factor), float(out.additional_attenuation_db)


def ir_min_phase(ir: np.ndarray) -> Optional[np.ndarray]:
    """convertToMinimumPhase for one channel (host-only); None where the reference keeps the linear-phase IR."""
    a = np.ascontiguousarray(ir, dtype=np.float64)
    out = np.zeros_like(a)
    st = capi.load().cpq_ir_min_phase(a.ctypes.data_as(_dp), a.size, out.ctypes.data_as(_dp))
    if st == capi.ERR_UNSUPPORTED:
        return None
    if st != capi.OK:
        raise capi.CpqError(st, "cpq_ir_min_phase")
    return out


def ir_decode_wav(data: bytes):
    """The reference's IR file decode (JUCE WavAudioFormat through float, then the input transform) -> (array [channels, frames],
    sample_rate, bits_per_sample, is_float)."""
    lib = capi.load()
    info = capi.IrFile()
    st = lib.cpq_ir_decode_wav(data, len(data), C.byref(info), None, 0)
    if st != capi.OK:
        raise capi.CpqError(st, "cpq_ir_decode_wav")
    out = np.zeros((info.channels, info.frames))
    st = lib.cpq_ir_decode_wav(data, len(data), C.byref(info), out.ctypes.data_as(_dp), out.size)
    if st != capi.OK:
        raise capi.CpqError(st, "cpq_ir_decode_wav")
    return out, float(info.sample_rate), int(info.bits_per_sample), bool(info.is_float)


def ir_trim_silence(ch0: np.ndarray, ch1: Optional[np.ndarray] = None) -> int:
    a = np.ascontiguousarray(ch0, dtype=np.float64)
    b = None if ch1 is None else np.ascontiguousarray(ch1, dtype=np.float64)
    return int(capi.load().cpq_ir_trim_silence(a.ctypes.data_as(_dp), b.ctypes.data_as(_dp) if b is not None else None, a.size))


def ir_mixed_phase(linear: np.ndarray, minimum: np.ndarray, sample_rate: float, lo_hz: float = 200.0, hi_hz: float = 1000.0):
    """convertToMixedPhaseFallback for one channel (host-only); None where the reference returns an empty buffer."""
    a = np.ascontiguousarray(linear, dtype=np.float64)
    b = np.ascontiguousarray(minimum, dtype=np.float64)
    out = np.zeros_like(a)
    st = capi.load().cpq_ir_mixed_phase(a.ctypes.data_as(_dp), b.ctypes.data_as(_dp), a.size, sample_rate, lo_hz, hi_hz, out.ctypes.data_as(_dp))
    if st == capi.ERR_UNSUPPORTED:
        return None
    if st != capi.OK:
        raise capi.CpqError(st, "cpq_ir_mixed_phase")
    return out


def ir_freq_peak_gain(ir_l: np.ndarray, ir_r: Optional[np.ndarray] = None) -> float:
    a = np.ascontiguousarray(ir_l, dtype=np.float64)
    b = None if ir_r is None else np.ascontiguousarray(ir_r, dtype=np.float64)
    return float(capi.load().cpq_ir_freq_peak_gain(a.ctypes.data_as(_dp), b.ctypes.data_as(_dp) if b is not None else None, a.size))


def ir_peak_latency(ir_l: np.ndarray, ir_r: Optional[np.ndarray] = None) -> int:
    """LoaderThread::estimatePeakLatencySamples (host-only helper of the C ABI)."""
    a = np.ascontiguousarray(ir_l, dtype=np.float64)
    b = None if ir_r is None else np.ascontiguousarray(ir_r, dtype=np.float64)
    return int(capi.load().cpq_ir_peak_latency(a.ctypes.data_as(_dp), b.ctypes.data_as(_dp) if b is not None else None, a.size))


def plan_layout_ex(ir_len: int, known_block: int, call_size: int, spec: Optional[capi.FilterSpec], n_callbacks: int):
    """Layer plan + gather plan for SetImpulse(known_block) with Add/Get calls of call_size samples (host-only):
    (layout, tail sources per tail layer, L0 ring source per callback, L0 ring count per callback)."""
    lib = capi.load()
    out = capi.Layout()
    src = (C.c_int64 * (2 * max(n_callbacks, 1)))()
    l0s = (C.c_int64 * max(n_callbacks, 1))()
    l0c = (C.c_int32 * max(n_callbacks, 1))()
    st = lib.cpq_plan_layout_ex(ir_len, known_block, call_size, C.byref(spec) if spec is not None else None, n_callbacks,
                                C.byref(out), src, l0s, l0c)
    if st != capi.OK:
        raise capi.CpqError(st, "plan_layout_ex")
    tails = [np.array(src[i * n_callbacks:(i + 1) * n_callbacks], dtype=np.int64) for i in range(max(out.num_layers - 1, 0))]
    return out, tails, np.array(l0s[:n_callbacks], dtype=np.int64), np.array(l0c[:n_callbacks], dtype=np.int32)


def plan_layout(ir_len: int, block_size: int, spec: Optional[capi.FilterSpec], n_callbacks: int):
    """Host-only layer plan + gather plan (no device needed)."""
    lib = capi.load()
    out = capi.Layout()
    src = (C.c_int64 * (2 * max(n_callbacks, 1)))()
    st = lib.cpq_plan_layout(ir_len, block_size, C.byref(spec) if spec is not None else None, n_callbacks,
                             C.byref(out), src)
    if st != capi.OK:
        raise capi.CpqError(st, "plan_layout")
    tails = [np.array(src[i * n_callbacks:(i + 1) * n_callbacks], dtype=np.int64) for i in range(max(out.num_layers - 1, 0))]
    return out, tails
