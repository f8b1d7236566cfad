"""cfg5 on one GPU, one profiled call (for `ncu --profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
from tests import signals
dev = torch.device("cuda", 0)
sr, T = 192000.0, 1920000 // 512 * 512
eng = ConvoPeqEngine(4, 2, sr, 512, T, device=0, conv_boundary=capi.CONV_OUTER, shared_ir=True)
spec = capi.default_filter_spec(sample_rate=sr)
for ch in range(2):
    eng.set_impulse(-1, ch, signals.synth_ir(2097152, 40 + ch), 1.0, spec)
for s in range(4):
    eng.set_eq(s, signals.to_band(signals.band_params(seed=7 + s)))
eng.set_epilogue(1.0, 0)
x = torch.randn(8, T, device=dev, dtype=torch.float64) * 0.1
io = x.clone()
for _ in range(2):
    io.copy_(x); torch.cuda.synchronize()
    eng.process_device(io.data_ptr(), T, T, capi.STAGE_ALL)
io.copy_(x); torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.process_device(io.data_ptr(), T, T, capi.STAGE_ALL)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
t = eng.timings()
print(f"total {t.total_ms:.3f} fwd {t.fft_fwd_ms:.3f} mac {t.mac_ms:.3f} inv {t.fft_inv_ms:.3f} eq {t.eq_ms:.3f}")
