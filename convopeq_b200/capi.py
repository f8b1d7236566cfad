"""ctypes binding of include/cpq.h (libcpq.so).  No torch types cross this boundary: pointers and sizes.

The library is built in-tree by convopeq_b200.build; importing this module never falls back to a CPU
implementation -- if libcpq.so is missing it is built, and if that fails the import raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from . import build as _build

NUM_BANDS = 20
MAX_LAYERS = 3

OK, ERR_INVALID, ERR_NOT_READY, ERR_CUDA, ERR_OOM, ERR_UNSUPPORTED, ERR_GEOMETRY = range(7)
STAGE_CONV, STAGE_EQ, STAGE_EPILOGUE, STAGE_ALL = 1, 2, 4, 7
STAGE_OUTPUT_FILTER, STAGE_FULL = 8, 15
STAGE_INPUT = 32
ORDER_EQ_THEN_CONV = 16
CONV_INNER, CONV_OUTER = 0, 1


class FilterSpec(C.Structure):
    _fields_ = [("sample_rate", C.c_double), ("hc_mode", C.c_int32), ("lc_mode", C.c_int32),
                ("tail_mode", C.c_int32), ("tail_enabled", C.c_int32), ("tail_start_seconds", C.c_double),
                ("tail_strength", C.c_double), ("tail_l1l2_multiplier", C.c_int32), ("reserved_", C.c_int32)]


class SvfCoeffs(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("a1", "a2", "a3", "m0", "m1", "m2")]


class EqBandParams(C.Structure):
    _fields_ = [("frequency", C.c_float), ("gain_db", C.c_float), ("q", C.c_float),
                ("enabled", C.c_int32), ("type", C.c_int32), ("channel_mode", C.c_int32)]


class IrScale(C.Structure):
    _fields_ = [("scale_factor", C.c_double), ("has_scale_factor", C.c_int32), ("additional_attenuation_db", C.c_float)]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_streams", C.c_int32), ("n_channels", C.c_int32),
                ("block_size", C.c_int32), ("sample_rate", C.c_double), ("max_samples", C.c_int64),
                ("conv_boundary", C.c_int32), ("shared_ir", C.c_int32), ("shared_eq", C.c_int32),
                ("uniform_partitions", C.c_int32), ("workspace_bytes", C.c_size_t)]


class LayerLayout(C.Structure):
    _fields_ = [("part_size", C.c_int32), ("fft_size", C.c_int32), ("num_parts_ir", C.c_int32),
                ("num_parts", C.c_int32), ("parts_per_callback", C.c_int32), ("output_delay_samples", C.c_int32),
                ("ir_offset", C.c_int32), ("ir_len", C.c_int32), ("first_output_sample", C.c_int64),
                ("skipped_callbacks", C.c_int64), ("gain", C.c_double)]


class Layout(C.Structure):
    _fields_ = [("num_layers", C.c_int32), ("latency", C.c_int32), ("layers", LayerLayout * MAX_LAYERS)]


class Timings(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("fft_fwd_ms", C.c_float), ("mac_ms", C.c_float), ("fft_inv_ms", C.c_float),
                ("eq_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float),
                ("kernel_launches", C.c_int32), ("chunks", C.c_int32)]


class IrFile(C.Structure):
    _fields_ = [("channels", C.c_int32), ("bits_per_sample", C.c_int32), ("is_float", C.c_int32), ("frames", C.c_int64),
                ("sample_rate", C.c_double)]


class IrLoadInfo(C.Structure):
    _fields_ = [("file_channels", C.c_int32), ("file_frames", C.c_int32), ("trimmed_frames", C.c_int32), ("target_length", C.c_int32),
                ("peak_latency", C.c_int32), ("phase_applied", C.c_int32), ("file_sample_rate", C.c_double), ("scale_factor", C.c_double)]


# every symbol include/cpq.h declares (tests check the library exports exactly these)
EXPORTS = [
    "cpq_abi_version", "cpq_status_string", "cpq_last_error", "cpq_filter_spec_default", "cpq_config_default",
    "cpq_create", "cpq_destroy", "cpq_reset", "cpq_set_impulse", "cpq_set_eq", "cpq_schedule_total_gain",
    "cpq_set_epilogue", "cpq_set_dither_uniforms", "cpq_set_output_filter", "cpq_set_output_stage", "cpq_set_conv_input_trim", "cpq_output_filter_design", "cpq_design_band", "cpq_db_to_gain", "cpq_equal_power_sin",
    "cpq_process", "cpq_process_f32", "cpq_process_device", "cpq_set_partition_range", "cpq_total_partitions", "cpq_get_layout",
    "cpq_latency", "cpq_get_timings", "cpq_get_eq_state", "cpq_cuda_stream", "cpq_kernel_launch_count",
    "cpq_plan_layout", "cpq_plan_layout_ex", "cpq_set_eq_mode", "cpq_band_node_active", "cpq_get_agc_state",
    "cpq_set_mix", "cpq_ir_peak_latency", "cpq_set_direct_head", "cpq_parse_eq_preset",
    "cpq_set_convolver_bypass", "cpq_set_peak_limiter", "cpq_set_input_gain", "cpq_set_partial_sources", "cpq_set_stream_window", "cpq_ir_scale_factor", "cpq_ir_freq_peak_gain", "cpq_ir_min_phase",
    "cpq_ir_target_length", "cpq_ir_prepare", "cpq_set_dither_seed", "cpq_set_dither_uniforms_device", "cpq_set_streaming", "cpq_stream_position", "cpq_state_size", "cpq_export_state", "cpq_import_state",
    "cpq_debug_check_guards", "cpq_probe_dfma_tflops", "cpq_probe_dfma_latency",
    "cpq_ir_decode_wav", "cpq_ir_trim_silence", "cpq_ir_mixed_phase", "cpq_load_impulse_wav",
]

_lib: Optional[C.CDLL] = None


def lib_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    # CPQ_LIB selects another build of the same library (tuning experiments: scripts/build_variant.py)
    path = os.environ.get("CPQ_LIB") or _build.build()   # no-op when up to date; raises when nvcc is unavailable and no .so exists
    if not os.path.exists(path):
        raise RuntimeError("libcpq.so missing: the CUDA extension is mandatory (no CPU fallback)")
    L = C.CDLL(path)
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    L.cpq_abi_version.restype = C.c_int
    L.cpq_status_string.restype = C.c_char_p
    L.cpq_status_string.argtypes = [C.c_int]
    L.cpq_last_error.restype = C.c_char_p
    L.cpq_last_error.argtypes = [vp]
    L.cpq_filter_spec_default.argtypes = [C.POINTER(FilterSpec)]
    L.cpq_filter_spec_default.restype = None
    L.cpq_config_default.argtypes = [C.POINTER(Config)]
    L.cpq_config_default.restype = None
    L.cpq_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.cpq_destroy.argtypes = [vp]
    L.cpq_destroy.restype = None
    L.cpq_reset.argtypes = [vp]
    L.cpq_set_impulse.argtypes = [vp, C.c_int, C.c_int, dp, C.c_int, C.c_double, C.POINTER(FilterSpec)]
    L.cpq_set_eq.argtypes = [vp, C.c_int, C.POINTER(SvfCoeffs), C.POINTER(C.c_uint8), C.POINTER(C.c_int32),
                             C.c_double, C.c_double]
    L.cpq_schedule_total_gain.argtypes = [vp, C.c_int, C.c_int64, C.c_double]
    L.cpq_set_eq_mode.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8)]
    L.cpq_band_node_active.argtypes = [C.c_int, C.c_float, C.c_int, C.c_double]
    L.cpq_get_agc_state.argtypes = [vp, C.c_int, dp]
    L.cpq_set_epilogue.argtypes = [vp, C.c_double, C.c_int]
    L.cpq_set_dither_uniforms.argtypes = [vp, dp, C.c_int64]
    L.cpq_set_output_filter.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.cpq_set_output_stage.argtypes = [vp, C.c_double, C.c_int]
    L.cpq_set_conv_input_trim.argtypes = [vp, C.c_double]
    L.cpq_set_mix.argtypes = [vp, C.c_float, C.c_int]
    L.cpq_set_direct_head.argtypes = [vp, C.c_int]
    L.cpq_set_convolver_bypass.argtypes = [vp, C.c_int]
    L.cpq_set_peak_limiter.argtypes = [vp, C.c_double]
    L.cpq_set_input_gain.argtypes = [vp, C.c_double]
    L.cpq_set_partial_sources.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.cpq_set_stream_window.argtypes = [vp, C.c_int, C.c_int]
    L.cpq_ir_scale_factor.argtypes = [dp, dp, C.c_int, dp, dp, C.c_int, C.c_double, C.POINTER(IrScale)]
    L.cpq_ir_min_phase.argtypes = [dp, C.c_int, dp]
    L.cpq_ir_freq_peak_gain.argtypes = [dp, dp, C.c_int]
    L.cpq_ir_freq_peak_gain.restype = C.c_double
    L.cpq_parse_eq_preset.argtypes = [C.c_char_p, C.POINTER(EqBandParams), C.POINTER(C.c_float)]
    L.cpq_ir_peak_latency.argtypes = [dp, dp, C.c_int]
    L.cpq_output_filter_design.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, dp]
    L.cpq_output_filter_design.restype = None
    L.cpq_design_band.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_double, C.POINTER(SvfCoeffs)]
    L.cpq_db_to_gain.argtypes = [C.c_float]
    L.cpq_db_to_gain.restype = C.c_double
    L.cpq_equal_power_sin.argtypes = [C.c_double]
    L.cpq_equal_power_sin.restype = C.c_double
    L.cpq_process.argtypes = [vp, C.POINTER(dp), C.c_int64, C.c_uint]
    L.cpq_process_f32.argtypes = [vp, C.POINTER(C.POINTER(C.c_float)), C.c_int64, C.c_uint]
    L.cpq_process_device.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_uint]
    L.cpq_set_partition_range.argtypes = [vp, C.c_int, C.c_int]
    L.cpq_total_partitions.argtypes = [vp]
    L.cpq_get_layout.argtypes = [vp, C.POINTER(Layout)]
    L.cpq_latency.argtypes = [vp]
    L.cpq_get_timings.argtypes = [vp, C.POINTER(Timings)]
    L.cpq_get_eq_state.argtypes = [vp, C.c_int, dp]
    L.cpq_cuda_stream.argtypes = [vp]
    L.cpq_cuda_stream.restype = vp
    L.cpq_kernel_launch_count.argtypes = [vp]
    L.cpq_kernel_launch_count.restype = C.c_int64
    L.cpq_plan_layout.argtypes = [C.c_int, C.c_int, C.POINTER(FilterSpec), C.c_int64, C.POINTER(Layout),
                                  C.POINTER(C.c_int64)]
    L.cpq_plan_layout_ex.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(FilterSpec), C.c_int64, C.POINTER(Layout),
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    L.cpq_ir_target_length.argtypes = [C.c_double, C.c_double]
    L.cpq_ir_prepare.argtypes = [dp, C.c_int, C.c_double, C.c_double, dp, C.c_int]
    L.cpq_set_dither_seed.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.cpq_set_dither_uniforms_device.argtypes = [vp, vp, C.c_int64]
    L.cpq_set_streaming.argtypes = [vp, C.c_int]
    L.cpq_stream_position.argtypes = [vp]
    L.cpq_stream_position.restype = C.c_int64
    L.cpq_state_size.argtypes = [vp]
    L.cpq_state_size.restype = C.c_size_t
    L.cpq_export_state.argtypes = [vp, vp, C.c_size_t]
    L.cpq_import_state.argtypes = [vp, vp, C.c_size_t]
    L.cpq_probe_dfma_tflops.argtypes = [C.c_int, C.c_int]
    L.cpq_probe_dfma_tflops.restype = C.c_double
    L.cpq_probe_dfma_latency.argtypes = [C.c_int]
    L.cpq_probe_dfma_latency.restype = C.c_double
    L.cpq_ir_decode_wav.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(IrFile), dp, C.c_size_t]
    L.cpq_ir_trim_silence.argtypes = [dp, dp, C.c_int]
    L.cpq_ir_mixed_phase.argtypes = [dp, dp, C.c_int, C.c_double, C.c_double, C.c_double, dp]
    L.cpq_load_impulse_wav.argtypes = [vp, C.c_int, C.c_char_p, C.c_size_t, C.c_int, C.c_double, C.POINTER(FilterSpec), C.POINTER(IrLoadInfo)]
    _lib = L
    return L


class CpqError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"cpq status {status}: {message}")
        self.status = status


def default_filter_spec(**kw) -> FilterSpec:
    s = FilterSpec()
    load().cpq_filter_spec_default(C.byref(s))
    for k, v in kw.items():
        setattr(s, k, v)
    return s
