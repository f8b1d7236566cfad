"""First-contact GPU script: prints per-stage errors against the checker (not a test, a debugging aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
from oracle.bindings import best_checker, Oracle, FilterSpec as OFilterSpec
from tests import signals

chk = best_checker(); orc = Oracle()
print("checker:", chk.kind)
L = capi.load()
print("dfma TFLOP/s:", L.cpq_probe_dfma_tflops(0, 20000))

def conv_case(ir_len, block, T, kw=None, sr=48000.0):
    ir = signals.synth_ir(ir_len, 2)
    x = np.stack([signals.noise(T, 1), signals.noise(T, 11)])
    eng = ConvoPeqEngine(1, 2, sr, block, T)
    cs = capi.default_filter_spec(**kw) if kw is not None else None
    eng.set_impulse(0, 0, ir, 1.0, cs); eng.set_impulse(0, 1, ir, 1.0, cs)
    y = x.copy(); eng.process(y, capi.STAGE_CONV)
    t = eng.timings()
    want, _ = chk.nuc_run(ir, x[0], block, spec=OFilterSpec(**kw) if kw is not None else None)
    err = np.abs(y[0] - want)
    print(f"conv ir={ir_len} B={block} T={T} spec={kw}: max err {err.max():.3e} at {err.argmax()} |y| {np.abs(want).max():.3f}"
          f"  fwd {t.fft_fwd_ms:.3f} mac {t.mac_ms:.3f} inv {t.fft_inv_ms:.3f} eq {t.eq_ms:.3f} ms")
    eng.close()

for c in [(512, 512, 4096), (4096, 512, 16384), (65536, 512, 65536), (65536, 512, 65536, {}), (65536, 64, 16384), (65536, 1024, 131072),
          (1000, 128, 4096), (3000, 256, 8192), (5000, 1024, 16384), (9000, 2048, 32768), (20000, 4096, 65536), (40000, 8192, 131072)]:
    try:
        conv_case(*c)
    except Exception as e:
        print("conv case", c, "failed:", e)

sr, block = 48000.0, 512
for T in (4096, 4096 * 3 + 512, 96000 // 512 * 512):
    for sat in (0.2, 0.0):
        params = signals.band_params(7)
        xl, xr = signals.log_sweep(T, sr)
        eng = ConvoPeqEngine(1, 2, sr, block, T)
        eng.set_eq(0, signals.to_band(params), sat, 0.0)
        y = np.stack([xl, xr]).copy(); eng.process(y, capi.STAGE_EQ)
        st = eng.eq_state(0); t = eng.timings(); eng.close()
        wl, wr, ws = chk.eq_run(signals.to_eqband(params), xl, xr, sr, block, saturation=sat)
        print(f"eq T={T} sat={sat}: errL {np.abs(y[0]-wl).max():.3e} errR {np.abs(y[1]-wr).max():.3e} state {np.abs(st-ws).max():.3e} eq {t.eq_ms:.3f} ms")
