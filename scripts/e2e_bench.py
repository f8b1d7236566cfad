"""The e2e leg of bench.py alone (cfg4 through cpq_process on a pinned host buffer): python scripts/e2e_bench.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
from tests import signals
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
S, T = 1024, bench.T_FULL
eng = ConvoPeqEngine(S, 2, 48000.0, 512, T, device=0, conv_boundary=capi.CONV_OUTER)
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
x = torch.randn(2 * S, T, device=dev, dtype=torch.float64, generator=g) * 0.1
host = torch.empty(2 * S, T, dtype=torch.float64).pin_memory()
stream = torch.cuda.ExternalStream(eng.cuda_stream(), device=dev)
ms = []
for i in range(steps + 1):
    host.copy_(x); torch.cuda.synchronize()
    t = bench._timed(stream, lambda: eng.process_host_ptrs(host.data_ptr(), T, T, capi.STAGE_ALL))
    if i: ms.append(t)
tm = eng.timings()
print(f"e2e {sum(ms)/len(ms):.2f} ms  ({2*S*T/ (sum(ms)/len(ms)*1e-3)/1e9:.3f} G ch-samples/s)  chunks {tm.chunks} h2d {tm.h2d_ms:.1f} d2h {tm.d2h_ms:.1f}")
