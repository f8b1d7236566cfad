"""Seeded random configurations of the whole path against the reference pieces: host block (power of two or not), IR length /
layer plan, FilterSpec, EQ structure / AGC / channel modes (incl. Mid/Side), saturation, processing order, mix, direct head,
output stages, several streams per handle with a small workspace (several sequence chunks)."""
import numpy as np
import pytest

from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine, ir_peak_latency
from oracle.bindings import FilterSpec as OFilterSpec
from tests import signals

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _known(block):
    k = 64
    while k < block:
        k *= 2
    return k


@pytest.mark.parametrize("seed", range(40))
def test_random_configuration(checker, oracle, seed):
    g = np.random.default_rng(1000 + seed)
    block = int(g.choice([64, 96, 100, 128, 256, 441, 480, 512, 960, 1000, 1024, 1125]))
    pow2 = (block & (block - 1)) == 0
    sr = float(g.choice([44100.0, 48000.0, 96000.0]))
    ir_len = int(g.choice([37, 3000, 20000, 70000, 140000]))
    n_cb = int(g.integers(10, 30)) * 2          # T must be even, whatever the block
    T = block * n_cb
    n = int(g.integers(1, 4))
    use_spec = bool(g.integers(0, 2))
    kw = dict(sample_rate=sr, tail_mode=int(g.integers(0, 3)), hc_mode=int(g.integers(0, 3)), lc_mode=int(g.integers(0, 2))) if use_spec else None
    ospec = OFilterSpec(**kw) if kw is not None else None
    cspec = capi.default_filter_spec(**kw) if kw is not None else None
    direct = bool(g.integers(0, 2)) and pow2            # the direct head needs a power-of-two host block here
    order_eq_first = bool(g.integers(0, 2))
    mix = float(g.choice([1.0, 1.0, 0.5, 0.0]))
    trim = float(g.choice([1.0, 0.6]))
    limiter = float(g.choice([0.0, 100.0]))
    makeup = float(g.choice([1.0, 3.0]))
    scale = float(g.choice([1.0, 0.5]))
    eng = ConvoPeqEngine(n, 2, sr, block, T, conv_boundary=capi.CONV_OUTER, workspace_bytes=int(g.choice([0, 30 << 20])))
    eng.set_direct_head(direct)
    x = np.stack([signals.noise(T, 5000 + 10 * seed + i, 0.3) for i in range(2 * n)])
    irs = [np.roll(signals.synth_ir(ir_len, 6000 + 10 * seed + i), int(g.integers(0, min(ir_len, 300)))) for i in range(2 * n)]
    bkws, ekws = [], []
    for s in range(n):
        agc = bool(g.integers(0, 2))
        structure = int(g.integers(0, 2))
        ms = bool(g.integers(0, 2))
        modes = [int(v) for v in g.integers(0, 5 if ms else 3, 20)]
        bkws.append(dict(seed=7000 + 10 * seed + s, modes=modes, types=[int(v) for v in g.integers(0, 5, 20)] if g.integers(0, 2) else None,
                         enabled=[int(v) for v in g.integers(0, 4, 20) > 0]))
        ekws.append(dict(agc=agc, structure=structure, saturation=float(g.choice([0.2, 0.0]))))
        for ch in range(2):
            eng.set_impulse(s, ch, irs[2 * s + ch], scale, cspec)
        eng.set_eq(s, signals.to_band(signals.band_params(**bkws[s])), ekws[s]["saturation"], 0.0, structure, agc)
    lat = 0 if direct else eng.latency()
    delay = lat + max(ir_peak_latency(irs[2 * s], irs[2 * s + 1]) for s in range(n))
    eng.set_mix(mix, delay)
    eng.set_conv_input_trim(trim)
    eng.set_output_filter(True, order_eq_first, 1, 0, 1)
    eng.set_output_stage(3.0, True)
    eng.set_peak_limiter(limiter)
    eng.set_epilogue(makeup, 0)
    y = x.copy()
    eng.process(y, capi.STAGE_FULL | (capi.ORDER_EQ_THEN_CONV if order_eq_first else 0))
    eng.close()

    def conv(v, s):
        out = []
        for ch in range(2):
            wet, _ = checker.nuc_run(irs[2 * s + ch], v[ch], _known(block), scale=scale, spec=ospec, call=block, direct_head=direct)
            out.append(oracle.outer_mix(wet, v[ch], mix, delay))
        return np.stack(out)

    def eq(v, s):
        l, r, _ = checker.eq_run(signals.to_eqband(signals.band_params(**bkws[s])), v[0], v[1], sr, block, **ekws[s])
        return np.stack([l, r])

    for s in range(n):
        v = x[2 * s:2 * s + 2]
        mid = conv(eq(v, s) * trim, s) if order_eq_first else eq(conv(v, s), s)
        want = checker.output_run(mid, sr, block, conv_is_last=order_eq_first, makeup=makeup, limiter_ms=limiter)
        err = np.abs(y[2 * s:2 * s + 2] - want).max()
        assert err <= TOL * max(1.0, np.abs(mid).max()), (seed, s, err, dict(block=block, sr=sr, ir_len=ir_len, spec=kw, direct=direct, eq_first=order_eq_first,
                                                                          mix=mix, eq=ekws[s], modes=bkws[s]["modes"]))


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration_mono_shared_and_gain_events(checker, oracle, seed):
    """Second family: mono or stereo handles, IR / EQ shared by all streams, a total-gain change at a random callback (50 ms ramp),
    conv -> EQ -> makeup + headroom only (no output stages), inner or outer convolver boundary."""
    g = np.random.default_rng(3000 + seed)
    block = int(g.choice([64, 128, 441, 480, 512, 2048]))
    sr = float(g.choice([48000.0, 96000.0]))
    ir_len = int(g.choice([500, 9000, 66000]))
    T = block * int(g.integers(10, 24)) * 2
    nch = int(g.choice([1, 2]))
    n = int(g.integers(1, 4))
    shared_ir, shared_eq = bool(g.integers(0, 2)), bool(g.integers(0, 2))
    outer = bool(g.integers(0, 2))
    gain_db = float(g.choice([0.0, -4.5]))
    change_at = int(g.integers(1, T // block - 1))
    change_db = float(g.choice([-9.0, 3.0]))
    sat = float(g.choice([0.2, 0.0]))
    makeup = float(g.choice([1.0, 0.7]))
    eng = ConvoPeqEngine(n, nch, sr, block, T, conv_boundary=capi.CONV_OUTER if outer else capi.CONV_INNER, shared_ir=shared_ir,
                         shared_eq=shared_eq, workspace_bytes=int(g.choice([0, 20 << 20])))
    x = np.stack([signals.noise(T, 8000 + 10 * seed + i, 0.3) for i in range(nch * n)])
    n_ir = nch if shared_ir else nch * n
    irs = [signals.synth_ir(ir_len, 8500 + 10 * seed + i) for i in range(n_ir)]
    n_eq = 1 if shared_eq else n
    bkws = [dict(seed=9000 + 10 * seed + s, modes=[int(v) for v in g.integers(0, 3, 20)]) for s in range(n_eq)]
    for s in range(1 if shared_ir else n):
        for ch in range(nch):
            eng.set_impulse(-1 if shared_ir else s, ch, irs[nch * s + ch], 1.0, None)
    for s in range(n_eq):
        eng.set_eq(-1 if shared_eq else s, signals.to_band(signals.band_params(**bkws[s])), sat, gain_db)
        eng.schedule_total_gain(-1 if shared_eq else s, change_at, change_db)
    eng.set_epilogue(makeup, 0)
    y = x.copy()
    eng.process(y, capi.STAGE_ALL)
    eng.close()
    known = _known(block)
    for s in range(n):
        chans = []
        for ch in range(nch):
            ir = irs[ch] if shared_ir else irs[nch * s + ch]
            wet, _ = checker.nuc_run(ir, x[nch * s + ch], known, call=block)
            chans.append(oracle.outer_wet(wet, 1.0) if outer else wet)
        bk = bkws[0] if shared_eq else bkws[s]
        l, r, _ = checker.eq_run(signals.to_eqband(signals.band_params(**bk)), chans[0], chans[1] if nch > 1 else None, sr, block,
                                 saturation=sat, total_gain_db=gain_db, gain_change_db=change_db, gain_change_at=change_at * block)
        outs = [l] if nch == 1 else [l, r]
        for ch in range(nch):
            want, _, _ = oracle.epilogue(outs[ch], makeup, sr, 0)
            err = np.abs(y[nch * s + ch] - want).max()
            assert err <= TOL, (seed, s, ch, err, dict(block=block, nch=nch, n=n, shared_ir=shared_ir, shared_eq=shared_eq, outer=outer))
