"""The dither stage alone on a small batch (for ncu): python scripts/dither_probe.py [--seed]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
dev = torch.device('cuda', 0)
S, T = 64, 65536
eng = ConvoPeqEngine(S, 2, 48000.0, 512, T, device=0)
g = torch.Generator(device=dev); g.manual_seed(1)
x = torch.randn(2 * S, T, device=dev, dtype=torch.float64, generator=g) * 0.1
eng.set_epilogue(1.0, 24)
if "--seed" in sys.argv:
    eng.set_dither_seed(list(range(1, S + 1)))
else:
    u = torch.rand(2 * S, 2 * T, device=dev, dtype=torch.float64, generator=g)
    eng.set_dither_uniforms_device(u.data_ptr(), T)
for i in range(3):
    io = x.clone(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.process_device(io.data_ptr(), T, T, capi.STAGE_EPILOGUE)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"epilogue + dither, {2*S} sequences x {T}: {dt*1e3:.2f} ms = {dt/T*1.9e9:.0f} cycles per sample at 1.9 GHz")
