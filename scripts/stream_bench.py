"""Streaming continuation at cfg4's geometry: time per call for calls of 1 / 8 / 64 / 938 callbacks (device-resident)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
from tests import signals
dev = torch.device('cuda', 0)
S, B = int(os.environ.get("STREAMS", "1024")), 512
eng = ConvoPeqEngine(S, 2, 48000.0, B, bench.T_FULL, device=0, conv_boundary=capi.CONV_OUTER)
g = torch.Generator(device=dev); g.manual_seed(1)
spec = capi.default_filter_spec()
decay = torch.exp(-torch.arange(131072, device=dev, dtype=torch.float64) / (131072 / 6.0)) / (131072 ** 0.5)
for s0 in range(0, 2 * S, 64):
    irs = (torch.randn(64, 131072, device=dev, dtype=torch.float64, generator=g) * decay).cpu().numpy()
    for i in range(64):
        eng.set_impulse((s0 + i) // 2, (s0 + i) % 2, irs[i], 1.0, spec)
for s in range(S):
    eng.set_eq(s, signals.to_band(signals.band_params(100 + s)), 0.2, 0.0)
eng.set_epilogue(1.0, 0)
eng.set_streaming(True)
for ncb in (1, 8, 64, 938):
    T = ncb * B
    x = torch.randn(2 * S, T, device=dev, dtype=torch.float64, generator=g) * 0.1
    eng.reset()
    calls = max(3, min(40, 2048 // ncb))
    ts = []
    for i in range(calls):
        io = x.clone(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.process_device(io.data_ptr(), T, T, capi.STAGE_ALL)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts = sorted(ts[1:])
    med = ts[len(ts) // 2]
    print(f"calls of {ncb:4d} callbacks ({T} samples x {2*S} sequences): median {med*1e3:8.3f} ms per call = {2*S*T/med/1e9:6.2f} G ch-samples/s "
          f"({med / (T / 48000.0) * 100:5.1f} % of real time)")
