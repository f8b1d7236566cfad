"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the two CPU checkers under oracle/.

* ``Oracle``  -> oracle/libcpq_oracle.so  (our C restatement, oracle/cpq_oracle.c; always buildable)
* ``Ref``     -> oracle/_ref/libcpq_ref.so (the reference's own TUs compiled in place; exists only when
  it was built where /root/reference is mounted -- the prebuilt .so travels to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Both classes expose the same small surface so a test can run either as "the reference".
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libcpq_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcpq_ref.so")
REFERENCE_ROOT = os.environ.get("CONVOPEQ_REF", "/root/reference")

_dp = C.POINTER(C.c_double)


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


class FilterSpec(C.Structure):
    """Mirror of convo::FilterSpec (MKLNonUniformConvolver.h:123-133), defaults included."""

    _fields_ = [
        ("sample_rate", C.c_double),
        ("hc_mode", C.c_int),
        ("lc_mode", C.c_int),
        ("tail_mode", C.c_int),
        ("tail_enabled", C.c_int),
        ("tail_start_seconds", C.c_double),
        ("tail_strength", C.c_double),
        ("tail_l1l2_multiplier", C.c_int),
    ]

    def __init__(self, sample_rate=48000.0, hc_mode=1, lc_mode=0, tail_mode=1, tail_enabled=1,
                 tail_start_seconds=0.085, tail_strength=1.0, tail_l1l2_multiplier=8):
        super().__init__(sample_rate, hc_mode, lc_mode, tail_mode, tail_enabled, tail_start_seconds,
                         tail_strength, tail_l1l2_multiplier)


class EqBand(C.Structure):
    _fields_ = [("frequency", C.c_float), ("gain_db", C.c_float), ("q", C.c_float),
                ("enabled", C.c_int), ("type", C.c_int), ("channel_mode", C.c_int)]


def node_active(b: "EqBand", sr: float) -> bool:
    """BandNode::active (createBandNode, EQProcessor.Coefficients.cpp:27-58): enabled, a prepared sample rate, and not a
    shelf/peaking band within 0.01 dB of flat.  Only the node path (taken when an active band is Mid/Side) uses it."""
    if not b.enabled or not sr > 0:
        return False
    if b.type not in (3, 4) and abs(np.float32(b.gain_db)) < np.float32(0.01):
        return False
    return True


def build(verbose: bool = False) -> None:
    """Compile the checkers (oracle always; _ref only where the reference tree is mounted)."""
    env = dict(os.environ, CONVOPEQ_REF=REFERENCE_ROOT)
    out = subprocess.run(["make", "-C", HERE, "all", f"CONVOPEQ_REF={REFERENCE_ROOT}"], env=env,
                         capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def have_oracle() -> bool:
    return os.path.exists(ORACLE_SO)


class _Base:
    prefix = ""
    lib: C.CDLL

    # ---- convolver -------------------------------------------------------------------------
    def nuc_run(self, ir: np.ndarray, x: np.ndarray, block: int, scale: float = 1.0,
                spec: Optional[FilterSpec] = None, call: Optional[int] = None, direct_head: bool = False, uniform: bool = False):
        """SetImpulse + (Add, Get) loop over x in calls of `call` (default = block) samples."""
        ir = np.ascontiguousarray(ir, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        h = self._nuc_create()
        try:
            if uniform and self.prefix != "cpqo_":
                raise ValueError("the uniform-partition extension is not a reference mode")
            ok = self._nuc_set_impulse(h, ir, block, scale, spec, int(direct_head) | (2 if uniform else 0))
            if not ok:
                raise RuntimeError("SetImpulse failed")
            layout = self._nuc_layout(h)
            self._f("nuc_process")(h, _p(x), _p(y), x.size, int(call or block))
        finally:
            self._f("nuc_destroy")(h)
        return y, layout

    def nuc_layout_only(self, ir_len: int, block: int, spec: Optional[FilterSpec] = None):
        ir = np.zeros(ir_len)
        ir[0] = 1.0
        h = self._nuc_create()
        try:
            if not self._nuc_set_impulse(h, ir, block, 1.0, spec):
                raise RuntimeError("SetImpulse failed")
            return self._nuc_layout(h)
        finally:
            self._f("nuc_destroy")(h)

    def nuc_spectra(self, ir: np.ndarray, block: int, scale: float = 1.0, spec: Optional[FilterSpec] = None):
        """Stored partition spectra per layer, list of complex arrays [numParts][P+1] in reference order."""
        ir = np.ascontiguousarray(ir, dtype=np.float64)
        h = self._nuc_create()
        try:
            if not self._nuc_set_impulse(h, ir, block, scale, spec):
                raise RuntimeError("SetImpulse failed")
            layout = self._nuc_layout(h)
            out = []
            for li, lay in enumerate(layout["layers"]):
                cs = lay["part_size"] + 1
                arr = np.zeros((lay["num_parts"], cs), dtype=np.complex128)
                re = np.zeros(cs)
                im = np.zeros(cs)
                for p in range(lay["num_parts"]):
                    n = self._f("nuc_ir_spectrum")(h, li, p, _p(re), _p(im))
                    assert n == cs
                    arr[p] = re + 1j * im
                out.append(arr)
            return out, layout
        finally:
            self._f("nuc_destroy")(h)

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _nuc_create(self):
        return self._f("nuc_create")()

    def _nuc_layout(self, h):
        lay = (C.c_int * 24)()
        g = (C.c_double * 3)()
        n = self._f("nuc_layout")(h, lay, g)
        keys = ["part_size", "num_parts_ir", "num_parts", "parts_per_callback", "output_delay_samples",
                "delay_capacity", "is_immediate", "fft_size"]
        return {"num_layers": n, "gains": [g[i] for i in range(3)],
                "layers": [dict(zip(keys, [lay[l * 8 + i] for i in range(8)])) for l in range(n)]}

    # ---- output stages ------------------------------------------------------------------------
    def output_design(self, sr: float, conv_is_last: bool, hc: int = 1, lc: int = 0, lp: int = 1) -> np.ndarray:
        """The three OutputFilter stages' {b0,b1,b2,a1,a2}, in processing order."""
        out = np.zeros(15)
        self._f("out_design")(sr, int(conv_is_last), hc, lc, lp, _p(out))
        return out.reshape(3, 5)

    def output_run(self, x: np.ndarray, sr: float, block: int, use_filter: bool = True, conv_is_last: bool = False, hc: int = 1,
                   lc: int = 0, lp: int = 1, makeup: float = 1.0, dc_cutoff: float = 3.0, headroom: bool = True,
                   clamp: bool = True, limiter_ms: float = 0.0) -> np.ndarray:
        """[OutputFilter] -> makeup -> [DC blocker] -> [headroom] -> [scrub + clamp] per callback on x[channels][T]."""
        y = np.ascontiguousarray(x, dtype=np.float64).copy()
        h = self._f("out_create")(sr, dc_cutoff if dc_cutoff > 0 else 1.0)
        if limiter_ms > 0:
            f = self._f("out_set_limiter")
            f.argtypes = [C.c_void_p, C.c_double]
            f.restype = None
            f(h, limiter_ms)
        try:
            self._f("out_process")(h, _p(y[0]), _p(y[1]) if y.shape[0] > 1 else None, y.shape[1], block, int(use_filter),
                                   int(conv_is_last), hc, lc, lp, makeup, int(dc_cutoff > 0), int(headroom), int(clamp))
        finally:
            self._f("out_destroy")(h)
        return y

    # ---- input stage -------------------------------------------------------------------------
    def input_transform(self, x: np.ndarray, gain: float = 1.0) -> np.ndarray:
        """convertDoubleToDoubleHighQuality: gain, NaN / denormal scrub, clamp to [-1, 1]."""
        d = np.ascontiguousarray(x, dtype=np.float64).copy()
        f = self._f("input_transform")
        f.argtypes = [_dp, C.c_long if self.prefix == "cpqo_" else C.c_int, C.c_double]
        f.restype = None
        f(_p(d), d.size, gain)
        return d

    # ---- IR preparation ----------------------------------------------------------------------
    def ir_freq_peak_gain(self, ir_l: np.ndarray, ir_r: Optional[np.ndarray] = None) -> float:
        """IRAnalyzer::estimateMaxFrequencyResponseGain."""
        a = np.ascontiguousarray(ir_l, dtype=np.float64)
        b = None if ir_r is None else np.ascontiguousarray(ir_r, dtype=np.float64)
        f = self._f("ir_freq_peak_gain")
        f.argtypes = [_dp, _dp, C.c_int]
        f.restype = C.c_double
        return float(f(_p(a), _p(b), a.size))

    def ir_dc_block(self, x: np.ndarray, sr: float, cutoff: float = 1.0) -> np.ndarray:
        """UltraHighRateDCBlocker::init(sr, cutoff) + process over one buffer (IR loader stage)."""
        d = np.ascontiguousarray(x, dtype=np.float64).copy()
        f = self._f("ir_dc_block")
        f.argtypes = [_dp, C.c_int, C.c_double, C.c_double]
        f.restype = None
        f(_p(d), d.size, sr, cutoff)
        return d

    # ---- EQ ----------------------------------------------------------------------------------
    def eq_design(self, type_: int, f: float, gain_db: float, q: float, sr: float) -> np.ndarray:
        out = np.zeros(6)
        self._f("eq_design")(int(type_), C.c_float(f), C.c_float(gain_db), C.c_float(q), C.c_double(sr), _p(out))
        return out


def _common_sigs(lib, pre, nuc_set_impulse_extra):
    vp = C.c_void_p
    f = lambda n: getattr(lib, pre + n)
    f("nuc_create").restype = vp
    f("nuc_destroy").argtypes = [vp]
    f("nuc_destroy").restype = None
    f("nuc_process").argtypes = [vp, _dp, _dp, C.c_long, C.c_int]
    f("nuc_process").restype = None
    f("nuc_layout").argtypes = [vp, C.POINTER(C.c_int), _dp]
    f("nuc_ir_spectrum").argtypes = [vp, C.c_int, C.c_int, _dp, _dp]
    f("eq_design").argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_double, _dp]
    f("eq_design").restype = None
    f("chain_process").argtypes = [vp, vp, vp, _dp, _dp, C.c_long, C.c_int, C.c_int, C.c_double, C.c_int]
    f("chain_process").restype = None
    f("eq_destroy").argtypes = [vp]
    f("eq_destroy").restype = None
    f("out_create").restype = vp
    f("out_create").argtypes = [C.c_double, C.c_double]
    f("out_destroy").argtypes = [vp]
    f("out_destroy").restype = None
    f("out_design").argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
    f("out_design").restype = None
    f("out_process").argtypes = [vp, _dp, _dp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                 C.c_int, C.c_int, C.c_int]
    f("out_process").restype = None


class Oracle(_Base):
    """Our C restatement (oracle/cpq_oracle.c)."""

    prefix = "cpqo_"
    kind = "port"

    def __init__(self):
        if not have_oracle():
            build()
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        _common_sigs(L, self.prefix, None)
        L.cpqo_nuc_set_impulse.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_double, C.POINTER(FilterSpec)]
        L.cpqo_nuc_set_impulse_ex.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(FilterSpec)]
        L.cpqo_eq_create.restype = C.c_void_p
        L.cpqo_eq_create.argtypes = [C.c_double, C.c_float]
        L.cpqo_eq_destroy.argtypes = [C.c_void_p]
        L.cpqo_eq_set_band.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_int]
        L.cpqo_eq_set_saturation.argtypes = [C.c_void_p, C.c_float]
        L.cpqo_eq_set_node_active.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.cpqo_eq_set_mode.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.cpqo_eq_get_agc.argtypes = [C.c_void_p, _dp]
        L.cpqo_eq_get_ms_state.argtypes = [C.c_void_p, _dp]
        L.cpqo_eq_set_total_gain.argtypes = [C.c_void_p, C.c_float]
        L.cpqo_eq_process.argtypes = [C.c_void_p, _dp, _dp, C.c_long, C.c_int]
        L.cpqo_eq_get_state.argtypes = [C.c_void_p, _dp]
        L.cpqo_epilogue.argtypes = [_dp, C.c_long, C.c_double, C.c_double, C.c_int, _dp, _dp, _dp]
        L.cpqo_epilogue.restype = None
        L.cpqo_epilogue_ex.argtypes = [_dp, C.c_long, C.c_double, C.c_double, C.c_int, _dp, _dp, _dp, C.c_int]
        L.cpqo_epilogue_ex.restype = None
        L.cpqo_outer_wet.argtypes = [_dp, C.c_long, C.c_double]
        L.cpqo_outer_wet.restype = None
        L.cpqo_outer_mix.argtypes = [_dp, _dp, C.c_long, C.c_float, C.c_int]
        L.cpqo_outer_mix.restype = None
        L.cpqo_ir_peak_latency.argtypes = [_dp, _dp, C.c_int]
        L.cpqo_ir_freq_peak_gain.argtypes = [_dp, _dp, C.c_int]
        L.cpqo_ir_freq_peak_gain.restype = C.c_double
        L.cpqo_ir_scale_factor.argtypes = [_dp, _dp, C.c_int, _dp, _dp, C.c_int, C.c_double, _dp]
        L.cpqo_ir_scale_factor.restype = None
        L.cpqo_equal_power_sin.argtypes = [C.c_double]
        L.cpqo_equal_power_sin.restype = C.c_double
        L.cpqo_db_to_gain.argtypes = [C.c_float]
        L.cpqo_db_to_gain.restype = C.c_double
        L.cpqo_dither_coeffs.argtypes = [C.c_double, C.c_int, _dp]

    def _nuc_set_impulse(self, h, ir, block, scale, spec, direct_head=False):
        return self.lib.cpqo_nuc_set_impulse_ex(h, _p(ir), ir.size, block, scale, int(direct_head),
                                                C.byref(spec) if spec is not None else None)

    def eq_run(self, bands: Sequence[EqBand], xl: np.ndarray, xr: Optional[np.ndarray], sr: float, block: int,
               saturation: float = 0.2, total_gain_db: float = 0.0, gain_change_db: Optional[float] = None,
               gain_change_at: int = 0, structure: int = 0, agc: bool = False):
        """createCoeffCache + process(block, params, cache) over the signal; returns (L, R, state[2][20][2])."""
        L = self.lib
        e = L.cpqo_eq_create(sr, C.c_float(total_gain_db))
        try:
            for i, b in enumerate(bands):
                active = bool(b.enabled) and sr > 0
                co = self.eq_design(b.type, b.frequency, b.gain_db, b.q, sr) if active else np.zeros(6)
                L.cpqo_eq_set_band(e, i, _p(co), int(active), int(b.channel_mode))
                L.cpqo_eq_set_node_active(e, i, int(node_active(b, sr)))
            L.cpqo_eq_set_saturation(e, C.c_float(saturation))
            L.cpqo_eq_set_mode(e, int(structure), int(agc))
            l = np.ascontiguousarray(xl, dtype=np.float64).copy()
            r = None if xr is None else np.ascontiguousarray(xr, dtype=np.float64).copy()
            if gain_change_db is None:
                rc = L.cpqo_eq_process(e, _p(l), _p(r), l.size, block)
            else:
                rc = L.cpqo_eq_process(e, _p(l[:gain_change_at]), _p(r[:gain_change_at]) if r is not None else None,
                                       gain_change_at, block)
                L.cpqo_eq_set_total_gain(e, C.c_float(gain_change_db))
                l2 = l[gain_change_at:]
                r2 = None if r is None else r[gain_change_at:]
                rc |= L.cpqo_eq_process(e, _p(l2), _p(r2), l2.size, block)
            if rc != 0:
                raise RuntimeError("EQ oracle: unsupported channel mode")
            st = np.zeros((2, 20, 2))
            L.cpqo_eq_get_state(e, _p(st))
        finally:
            L.cpqo_eq_destroy(e)
        return l, r, st

    def epilogue(self, x: np.ndarray, makeup_gain: float, sr: float, bit_depth: int,
                 uniforms: Optional[np.ndarray] = None, role: int = 1, z: Optional[np.ndarray] = None):
        """One channel of makeup + headroom / dither.  role 0 = left channel of a stereo block, 1 = right channel or mono
        (the two associations the compiled reference uses, see cpqo_epilogue_ex); z = carried error history (in/out)."""
        d = np.ascontiguousarray(x, dtype=np.float64).copy()
        z = np.zeros(12) if z is None else z
        tmp = np.zeros_like(d)
        self.lib.cpqo_epilogue_ex(_p(d), d.size, makeup_gain, sr, bit_depth, _p(uniforms), _p(z), _p(tmp), int(role))
        return d, tmp, z

    def dither_run(self, x: np.ndarray, uniforms: np.ndarray, sr: float, bit_depth: int, block: int = 512):
        """Same call as Ref.dither_run through the restatement: x[1 or 2][T], uniforms[channels][2T] -> (quantised, z[ch][12])."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        zs = np.zeros((x.shape[0], 12))
        for ch in range(x.shape[0]):
            role = 0 if (x.shape[0] == 2 and ch == 0) else 1
            out[ch], _, zs[ch] = self.epilogue(x[ch], 1.0, sr, bit_depth, np.ascontiguousarray(uniforms[ch]), role=role)
        return out, zs

    def dither_run_seeded(self, x: np.ndarray, seed: int, sr: float, bit_depth: int, block: int = 512):
        """PsychoacousticDither(seed) with its VSL stream unavailable (the header's own xorshift64* fallback generator)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = np.zeros((x.shape[0], 2 * x.shape[1]))
        f = self.lib.cpqo_dither_fallback_uniforms
        f.argtypes = [C.c_uint64, C.c_int, C.c_long, _dp]
        f.restype = None
        for ch in range(x.shape[0]):
            f(C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), ch, 2 * x.shape[1], _p(u[ch]))
        return self.dither_run(x, u, sr, bit_depth, block)

    def outer_mix(self, wet: np.ndarray, dry_in: np.ndarray, mix: float, delay: int) -> np.ndarray:
        """ConvolverProcessor::process, settled: scrub(wet) * sin-gain(mix) + delayed dry * sin-gain(1 - mix) (restated, unpinned)."""
        d = np.ascontiguousarray(wet, dtype=np.float64).copy()
        x = np.ascontiguousarray(dry_in, dtype=np.float64)
        self.lib.cpqo_outer_mix(_p(d), _p(x), d.size, C.c_float(mix), int(delay))
        return d

    def ir_peak_latency(self, ir_l: np.ndarray, ir_r: Optional[np.ndarray] = None) -> int:
        a = np.ascontiguousarray(ir_l, dtype=np.float64)
        b = None if ir_r is None else np.ascontiguousarray(ir_r, dtype=np.float64)
        return int(self.lib.cpqo_ir_peak_latency(_p(a), _p(b), a.size))

    def ir_scale_factor(self, ir_l, ir_r=None, cur_l=None, cur_r=None, cur_scale: float = 1.0):
        """IRConverter::computeScaleFactor -> (scaleFactor, hasScaleFactor, additionalAttenuationDb)."""
        a = np.ascontiguousarray(ir_l, dtype=np.float64)
        b = None if ir_r is None else np.ascontiguousarray(ir_r, dtype=np.float64)
        c = None if cur_l is None else np.ascontiguousarray(cur_l, dtype=np.float64)
        d = None if cur_r is None else np.ascontiguousarray(cur_r, dtype=np.float64)
        out = np.zeros(3)
        self.lib.cpqo_ir_scale_factor(_p(a), _p(b), a.size, _p(c), _p(d), 0 if c is None else c.size, cur_scale, _p(out))
        return float(out[0]), bool(out[1]), float(out[2])

    def ir_prepare(self, ir: np.ndarray, sr: float, target_seconds: float) -> np.ndarray:
        """DC blocker -> asymmetric Tukey -> trim to targetLength with fade-out (restated; the DC stage is pinned)."""
        a = np.ascontiguousarray(ir, dtype=np.float64)
        n = min(max(int(sr * float(np.float32(target_seconds))), 1), 2097152)
        out = np.zeros(n)
        f = self.lib.cpqo_ir_prepare
        f.argtypes = [_dp, C.c_int, C.c_double, C.c_double, _dp]
        f.restype = C.c_int
        got = f(_p(a), a.size, sr, target_seconds, _p(out))
        assert got == n
        return out

    def outer_wet(self, x: np.ndarray, mix: float = 1.0) -> np.ndarray:
        d = np.ascontiguousarray(x, dtype=np.float64).copy()
        self.lib.cpqo_outer_wet(_p(d), d.size, mix)
        return d


class Ref(_Base):
    """The reference's own code (oracle/_ref/libcpq_ref.so)."""

    prefix = "cpqref_"
    kind = "reference"

    def __init__(self):
        if not have_ref():
            raise FileNotFoundError(REF_SO)
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        _common_sigs(L, self.prefix, None)
        L.cpqref_nuc_set_impulse.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(FilterSpec)]
        L.cpqref_eq_create.restype = C.c_void_p
        L.cpqref_eq_create.argtypes = [C.c_double, C.c_int, C.c_float]
        L.cpqref_eq_destroy.argtypes = [C.c_void_p]
        L.cpqref_eq_set_params.argtypes = [C.c_void_p, C.POINTER(EqBand), C.c_float, C.c_int, C.c_int]
        L.cpqref_eq_set_total_gain.argtypes = [C.c_void_p, C.c_float]
        L.cpqref_eq_process.argtypes = [C.c_void_p, _dp, _dp, C.c_long, C.c_int]
        L.cpqref_eq_get_state.argtypes = [C.c_void_p, _dp]

    def dither_run(self, x: np.ndarray, uniforms: np.ndarray, sr: float, bit_depth: int, block: int = 512,
                   headroom: float = 0.8912509381337456):
        """The reference's own PsychoacousticDither::processStereoBlock (PsychoacousticDither.h:293-405, compiled in place) on
        x[1 or 2][T] with the VSL uniforms replaced by `uniforms`[channels][2T]; returns (quantised, shaper state [ch][12])."""
        d = np.ascontiguousarray(x, dtype=np.float64).copy()
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        assert u.shape == (d.shape[0], 2 * d.shape[1])
        z = np.zeros((2, 12))
        f = self.lib.cpqref_dither_process
        f.argtypes = [_dp, _dp, C.c_long, C.c_int, C.c_double, C.c_int, C.c_double, _dp, _dp, _dp]
        f.restype = None
        f(_p(d[0]), _p(d[1]) if d.shape[0] > 1 else None, d.shape[1], block, sr, bit_depth, headroom,
          _p(u[0]), _p(u[1]) if d.shape[0] > 1 else None, _p(z))
        return d, z[:d.shape[0]]

    def dither_run_seeded(self, x: np.ndarray, seed: int, sr: float, bit_depth: int, block: int = 512,
                          headroom: float = 0.8912509381337456):
        """The reference's PsychoacousticDither(seed) with vslNewStream failing: its own fallback generator supplies the uniforms."""
        d = np.ascontiguousarray(x, dtype=np.float64).copy()
        z = np.zeros((2, 12))
        f = self.lib.cpqref_dither_process_fallback
        f.argtypes = [_dp, _dp, C.c_long, C.c_int, C.c_double, C.c_int, C.c_double, C.c_uint64, _dp]
        f.restype = None
        f(_p(d[0]), _p(d[1]) if d.shape[0] > 1 else None, d.shape[1], block, sr, bit_depth, headroom, C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), _p(z))
        return d, z[:d.shape[0]]

    def _nuc_set_impulse(self, h, ir, block, scale, spec, direct_head=False):
        return self.lib.cpqref_nuc_set_impulse(h, _p(ir), ir.size, block, scale, int(direct_head),
                                               C.byref(spec) if spec is not None else None)

    def eq_run(self, bands: Sequence[EqBand], xl, xr, sr, block, saturation=0.2, total_gain_db=0.0,
               gain_change_db=None, gain_change_at=0, structure=0, agc=False):
        L = self.lib
        e = L.cpqref_eq_create(sr, max(block, 1), C.c_float(total_gain_db))
        try:
            arr = (EqBand * 20)(*bands)
            if not L.cpqref_eq_set_params(e, arr, C.c_float(saturation), int(structure), int(agc)):
                raise RuntimeError("createCoeffCache failed")
            l = np.ascontiguousarray(xl, dtype=np.float64).copy()
            r = None if xr is None else np.ascontiguousarray(xr, dtype=np.float64).copy()
            if gain_change_db is None:
                L.cpqref_eq_process(e, _p(l), _p(r), l.size, block)
            else:
                L.cpqref_eq_process(e, _p(l[:gain_change_at]), _p(r[:gain_change_at]) if r is not None else None,
                                    gain_change_at, block)
                L.cpqref_eq_set_total_gain(e, C.c_float(gain_change_db))
                l2 = l[gain_change_at:]
                r2 = None if r is None else r[gain_change_at:]
                L.cpqref_eq_process(e, _p(l2), _p(r2), l2.size, block)
            st = np.zeros((2, 20, 2))
            L.cpqref_eq_get_state(e, _p(st))
        finally:
            L.cpqref_eq_destroy(e)
        return l, r, st


def _chain_run(self, irs, bands, x, sr, block, spec, saturation, total_gain_db, makeup, outer, do_eq, do_epilogue,
               structure=0, agc=False, known_block=None):
    """One stream through the per-callback ConvolverThenEQ chain (conv -> wet gain -> EQ -> makeup*headroom).
    irs: (irL, irR) or None; x: [2, T] (copied). Returns y [2, T]. Releases the GIL inside the C call."""
    L = self.lib
    pre = self.prefix
    y = np.ascontiguousarray(x, dtype=np.float64).copy()
    nucs = [None, None]
    eq = None
    try:
        if irs is not None:
            for c in range(2):
                nucs[c] = self._nuc_create()
                # the application prepares the convolver with the host block rounded up to a power of two (knownBlockSize) and
                # calls it with the host block (preferredCallSize), LoaderThread.cpp:230,239-245
                if not self._nuc_set_impulse(nucs[c], np.ascontiguousarray(irs[c], dtype=np.float64), known_block or block, 1.0, spec):
                    raise RuntimeError("SetImpulse failed")
        if do_eq:
            if pre == "cpqref_":
                eq = L.cpqref_eq_create(sr, block, C.c_float(total_gain_db))
                arr = (EqBand * 20)(*bands)
                L.cpqref_eq_set_params(eq, arr, C.c_float(saturation), int(structure), int(agc))
            else:
                eq = L.cpqo_eq_create(sr, C.c_float(total_gain_db))
                for i, b in enumerate(bands):
                    active = bool(b.enabled) and sr > 0
                    co = self.eq_design(b.type, b.frequency, b.gain_db, b.q, sr) if active else np.zeros(6)
                    L.cpqo_eq_set_band(eq, i, _p(co), int(active), int(b.channel_mode))
                    L.cpqo_eq_set_node_active(eq, i, int(node_active(b, sr)))
                L.cpqo_eq_set_saturation(eq, C.c_float(saturation))
                L.cpqo_eq_set_mode(eq, int(structure), int(agc))
        self._f("chain_process")(nucs[0], nucs[1], eq, _p(y[0]), _p(y[1]), y.shape[1], block, int(outer),
                                 float(makeup), int(do_epilogue))
    finally:
        for n in nucs:
            if n:
                self._f("nuc_destroy")(n)
        if eq:
            self._f("eq_destroy")(eq)
    return y


def _chain_prepare(self, irs, bands, sr, block, spec, saturation=0.2, total_gain_db=0.0):
    """Build (nucL, nucR, eq) once so a benchmark can time processing only."""
    L = self.lib
    nucs = []
    for c in range(2):
        n = self._nuc_create()
        if not self._nuc_set_impulse(n, np.ascontiguousarray(irs[c], dtype=np.float64), block, 1.0, spec):
            raise RuntimeError("SetImpulse failed")
        nucs.append(n)
    if self.prefix == "cpqref_":
        eq = L.cpqref_eq_create(sr, block, C.c_float(total_gain_db))
        arr = (EqBand * 20)(*bands)
        L.cpqref_eq_set_params(eq, arr, C.c_float(saturation), 0, 0)
    else:
        eq = L.cpqo_eq_create(sr, C.c_float(total_gain_db))
        for i, b in enumerate(bands):
            active = bool(b.enabled) and sr > 0
            co = self.eq_design(b.type, b.frequency, b.gain_db, b.q, sr) if active else np.zeros(6)
            L.cpqo_eq_set_band(eq, i, _p(co), int(active), int(b.channel_mode))
        L.cpqo_eq_set_saturation(eq, C.c_float(saturation))
    return nucs[0], nucs[1], eq


def _chain_process_prepared(self, handles, y, block, outer=1, makeup=1.0, epilogue=1):
    self._f("chain_process")(handles[0], handles[1], handles[2], _p(y[0]), _p(y[1]), y.shape[1], block, int(outer),
                             float(makeup), int(epilogue))


def _chain_free(self, handles):
    self._f("nuc_destroy")(handles[0])
    self._f("nuc_destroy")(handles[1])
    self._f("eq_destroy")(handles[2])


def _install_chain():
    for cls in (Oracle, Ref):
        def chain_run(self, irs, bands, x, sr, block, spec=None, saturation=0.2, total_gain_db=0.0, makeup=1.0,
                      outer=True, do_eq=True, do_epilogue=True, structure=0, agc=False, known_block=None):
            return _chain_run(self, irs, bands, x, sr, block, spec, saturation, total_gain_db, makeup, outer, do_eq, do_epilogue,
                              structure, agc, known_block)
        cls.chain_run = chain_run
        cls.chain_prepare = _chain_prepare
        cls.chain_process_prepared = _chain_process_prepared
        cls.chain_free = _chain_free


def best_checker():
    """The strongest checker available: the real reference if it was compiled, else the restatement."""
    return Ref() if have_ref() else Oracle()


_install_chain()
