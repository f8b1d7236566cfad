"""torchrun check of the two multi-GPU modes on real GPUs (NCCL over NVLink):
  1. stream sharding (no collective): every rank runs its stream range, one stream per rank is checked against the CPU checker
  2. partition-range sharding (cfg 5 shape): partials summed with NCCL, then EQ + epilogue, checked on rank 0
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
from convopeq_b200.dist import stream_range, process_partition_sharded
from oracle.bindings import best_checker, Oracle, FilterSpec as OFilterSpec
from tests import signals

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
chk, orc = best_checker(), Oracle()

# ---- 1. stream sharding ----
sr, block, T, ir_len, n_streams = 48000.0, 512, 32768, 131072, 4 * world
b, e = stream_range(n_streams, rank, world)
eng = ConvoPeqEngine(e - b, 2, sr, block, T, device=local, conv_boundary=capi.CONV_OUTER)
x = np.stack([signals.noise(T, 100 + 2 * s + ch) for s in range(b, e) for ch in range(2)])
for i, s in enumerate(range(b, e)):
    for ch in range(2):
        eng.set_impulse(i, ch, signals.synth_ir(ir_len, 500 + 2 * s + ch), 1.0, capi.default_filter_spec())
    eng.set_eq(i, signals.to_band(signals.band_params(900 + s)))
eng.set_epilogue(1.0, 0)
y = x.copy()
eng.process(y, capi.STAGE_ALL)
eng.close()
s = b   # check this rank's first stream
want = chk.chain_run((signals.synth_ir(ir_len, 500 + 2 * s), signals.synth_ir(ir_len, 501 + 2 * s)), signals.to_eqband(signals.band_params(900 + s)),
                     x[0:2], sr, block, OFilterSpec())
err1 = float(np.abs(y[0:2] - want).max())

# ---- 2. partition-range sharding of one long IR (8 channels = 4 stereo pairs) ----
sr5, T5, ir5 = 192000.0, 131072, 2097152
eng = ConvoPeqEngine(4, 2, sr5, block, T5, device=local, conv_boundary=capi.CONV_OUTER, shared_ir=True, shared_eq=True)
irs = [signals.synth_ir(ir5, 40), signals.synth_ir(ir5, 41)]
spec = capi.default_filter_spec(sample_rate=sr5)
for ch in range(2):
    eng.set_impulse(-1, ch, irs[ch], 1.0, spec)
bands = signals.band_params(77)
eng.set_eq(-1, signals.to_band(bands))
eng.set_epilogue(1.0, 0)
x5 = np.stack([signals.noise(T5, 700 + i) for i in range(8)])
io = torch.from_numpy(x5).to(dev)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
process_partition_sharded(eng, io, T5, rank, world)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
total_parts = eng.total_partitions()
eng.close()
err2 = -1.0
if rank == 0:
    got = io[0:2].cpu().numpy()
    want = chk.chain_run(irs, signals.to_eqband(bands), x5[0:2], sr5, block, OFilterSpec(sample_rate=sr5))
    err2 = float(np.abs(got - want).max())
errs = torch.tensor([err1, err2], device=dev, dtype=torch.float64)
dist.all_reduce(errs, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"multi_gpu_check world={world}: stream-sharded max err {errs[0].item():.3e}; partition-sharded ({total_parts} partitions over {world} ranks, "
          f"NCCL all-reduce of 8x{T5} partials) max err {errs[1].item():.3e}, {dt*1e3:.1f} ms; checker={chk.kind}")
    assert errs[0].item() <= 1e-10 and errs[1].item() <= 1e-10
dist.destroy_process_group()
