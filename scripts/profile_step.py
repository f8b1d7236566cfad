"""One step of the cfg4-shaped hot path between cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
from tests import signals

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=512)
ap.add_argument("--samples", type=int, default=65536)
ap.add_argument("--ir-len", type=int, default=131072)
ap.add_argument("--block", type=int, default=512)
ap.add_argument("--sr", type=float, default=48000.0)
ap.add_argument("--stages", type=int, default=7)
ap.add_argument("--warm", type=int, default=2)
ap.add_argument("--workspace-mb", type=int, default=0)
a = ap.parse_args()
S, T = a.streams, a.samples // a.block * a.block
dev = torch.device("cuda", 0)
eng = ConvoPeqEngine(S, 2, a.sr, a.block, T, 0, capi.CONV_OUTER, workspace_bytes=a.workspace_mb << 20)
g = torch.Generator(device=dev); g.manual_seed(1)
decay = torch.exp(-torch.arange(a.ir_len, device=dev, dtype=torch.float64) / (a.ir_len / 6.0)) / (a.ir_len ** 0.5)
spec = capi.default_filter_spec(sample_rate=a.sr)
for s0 in range(0, 2 * S, 64):
    n = min(64, 2 * S - s0)
    irs = (torch.randn(n, a.ir_len, device=dev, dtype=torch.float64, generator=g) * decay).cpu().numpy()
    for i in range(n):
        eng.set_impulse((s0 + i) // 2, (s0 + i) % 2, irs[i], 1.0, spec)
for s in range(S):
    eng.set_eq(s, signals.to_band(signals.band_params(100 + s)), 0.2, 0.0)
eng.set_epilogue(1.0, 0)
x = torch.randn(2 * S, T, device=dev, dtype=torch.float64, generator=g) * 0.1
io = torch.empty_like(x)
for _ in range(a.warm):
    io.copy_(x); torch.cuda.synchronize()
    eng.process_device(io.data_ptr(), T, T, a.stages)
io.copy_(x); torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.process_device(io.data_ptr(), T, T, a.stages)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
t = eng.timings()
print(f"step: total {t.total_ms:.3f} ms fwd {t.fft_fwd_ms:.3f} mac {t.mac_ms:.3f} inv {t.fft_inv_ms:.3f} eq {t.eq_ms:.3f} chunks {t.chunks} launches {t.kernel_launches}")
