"""torchrun check of partition-range sharding with the reduce fused into the EQ launch's load: every rank's
partial lives in torch symmetric memory, peers read it over NVLink while loading their tiles (cpq_set_partial_sources),
each rank finishes only its own streams (cpq_set_stream_window).  Compared with the NCCL all-reduce path and the CPU checker.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 scripts/fused_reduce_check.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.makedirs("gpurun_out", exist_ok=True)
LOG = open(f"gpurun_out/fused_rank{rank}.log", "w")
def log(*a):
    print(f"[{rank}]", *a, file=LOG, flush=True)

torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
log("pg up")
from convopeq_b200 import capi
from convopeq_b200.engine import ConvoPeqEngine
from convopeq_b200.dist import stream_range, process_partition_sharded, process_partition_sharded_fused
from oracle.bindings import best_checker, FilterSpec as OFilterSpec
from tests import signals
import torch.distributed._symmetric_memory as symm_mem

sr, block = 192000.0, 512
T = int(os.environ.get("CPQ_T", 65536))
ir_len = int(os.environ.get("CPQ_IR", 2097152))
n_streams = 4
irs = [signals.synth_ir(ir_len, 40), signals.synth_ir(ir_len, 41)]
spec = capi.default_filter_spec(sample_rate=sr)
bands = signals.band_params(77)
x = np.stack([signals.noise(T, 700 + i) for i in range(2 * n_streams)])

def make_engine():
    eng = ConvoPeqEngine(n_streams, 2, sr, block, T, device=local, conv_boundary=capi.CONV_OUTER, shared_ir=True, shared_eq=True)
    for ch in range(2):
        eng.set_impulse(-1, ch, irs[ch], 1.0, spec)
    eng.set_eq(-1, signals.to_band(bands))
    eng.set_epilogue(1.0, 0)
    return eng

eng = make_engine()
log("engine ready")
buf = symm_mem.empty((2 * n_streams, T), dtype=torch.float64, device=dev)
log("symm empty")
hdl = symm_mem.rendezvous(buf, dist.group.WORLD.group_name)
ptrs = [int(p) for p in hdl.buffer_ptrs]
log("rendezvous", [hex(p) for p in ptrs])
first, end = stream_range(n_streams, rank, world)
count = end - first

def bar():
    torch.cuda.synchronize()
    dist.barrier()

def run_fused():
    buf.copy_(torch.from_numpy(x))
    bar()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    if count > 0:
        process_partition_sharded_fused(eng, buf, ptrs, T, rank, world, (first, count), bar)
    else:   # more ranks than streams: contribute the partial only
        from convopeq_b200.dist import partition_ranges
        lay = eng.layout()
        b_, e_ = partition_ranges([lay.layers[i].num_parts_ir for i in range(lay.num_layers)], world)[rank]
        eng.set_partition_range(b_, e_)
        eng.process_device(buf.data_ptr(), T, T, capi.STAGE_CONV)
        bar(); bar()
        eng.set_partition_range(0, -1)
    torch.cuda.synchronize()
    return time.perf_counter() - t0

def run_nccl():
    io = torch.from_numpy(x).to(dev)
    bar()
    t0 = time.perf_counter()
    process_partition_sharded(eng, io, T, rank, world)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, io

run_fused(); log("fused warm-up done")
tf = min(run_fused() for _ in range(3)); log("fused", tf)
got = buf[2 * first:2 * end].cpu().numpy() if count > 0 else None
run_nccl(); tn, io = min((run_nccl() for _ in range(3)), key=lambda r: r[0]); log("nccl", tn)
err_f = err_n = 0.0
if count > 0:
    chk = best_checker()
    for s in range(first, end):
        want = chk.chain_run(irs, signals.to_eqband(bands), x[2 * s:2 * s + 2], sr, block, OFilterSpec(sample_rate=sr))
        err_f = max(err_f, float(np.abs(got[2 * (s - first):2 * (s - first) + 2] - want).max()))
        err_n = max(err_n, float(np.abs(io[2 * s:2 * s + 2].cpu().numpy() - want).max()))
t = torch.tensor([err_f, err_n, tf, tn], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    msg = (f"fused_reduce_check world={world}: {2 * n_streams} channels x {T} samples, {ir_len}-tap IR, {eng.total_partitions()} partitions; "
           f"fused (partials summed in the EQ load over NVLink peer memory, each rank finishes its own streams) {t[2].item()*1e3:.1f} ms, max err {t[0].item():.3e}; "
           f"NCCL all-reduce + EQ of every stream on every rank {t[3].item()*1e3:.1f} ms, max err {t[1].item():.3e}")
    print(msg)
    open("gpurun_out/fused_reduce_check.txt", "w").write(msg + "\n")
    assert t[0].item() <= 1e-10 and t[1].item() <= 1e-10
eng.close()
dist.destroy_process_group()
