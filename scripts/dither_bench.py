import sys, os
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import bench
from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
from tests import signals
dev = torch.device('cuda', 0)
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
x = torch.randn(2 * S, T, device=dev, dtype=torch.float64, generator=g) * 0.1
u = torch.rand(2 * S, 2 * T, device=dev, dtype=torch.float64, generator=g)
eng.set_epilogue(1.0, 24)
eng.set_dither_uniforms_device(u.data_ptr(), T)
print('dither24', bench.measure_workload(eng, x, T, capi.STAGE_ALL, steps=2, warm=1, host=False))
eng.set_epilogue(1.0, 0)
print('plain', bench.measure_workload(eng, x, T, capi.STAGE_ALL, steps=2, warm=1, host=False))
# dither alone
eng.set_epilogue(1.0, 24)
print('dither only', bench.measure_workload(eng, x, T, capi.STAGE_EPILOGUE, steps=2, warm=1, host=False))
