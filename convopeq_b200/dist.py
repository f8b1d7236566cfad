"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on B200, gloo in CPU tests).

Two ways the hot path shards:

* by stream -- every (stream, channel) has its own convolver state and every stream its own EQ state
  (ConvolverProcessor.h:669, EQProcessor.h:637), so ranks own contiguous stream ranges and there is NO
  data-path collective.  `stream_range` is all that is needed.
* by partition range, for one very long IR (BASELINE config 5) -- the convolver is linear in the IR, so each
  rank multiply-accumulates only its share of the flattened (layer, partition) list and produces a partial
  time-domain signal; one sum over ranks (NCCL reduce-scatter by channel pair over NVLink, or all-reduce)
  yields the wet signal, after which EQ + epilogue run on the owning rank.  EQ and dither are sequential in
  time per channel and are never time-sharded.

Nothing here touches the arithmetic: the kernels are the same single-GPU kernels.
"""
from __future__ import annotations

from typing import List, Tuple


def stream_range(n_streams: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of streams owned by `rank`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_streams, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def bind_to_gpu_numa_node(device_index: int, local_rank: int = 0, local_world: int = 1):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, BEFORE it allocates pinned host buffers (first touch
    then places them next to the GPU's PCIe root).  With one process per GPU and host <-> device streaming on every rank,
    buffers on the far socket halve the transfer rate.  When several local ranks would end up on the same CPU list (boxes
    that expose one NUMA node for every GPU) each rank takes its own slice of it, so the ranks' copy threads do not share
    cores.  Returns the cpu list it bound to, or None when the topology is not exposed (then nothing changes)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:      # nvml prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        cpus = sorted(cpus)
        if local_world > 1:
            # which local GPUs share this list?  slice it among them
            same = []
            for d in range(pynvml.nvmlDeviceGetCount()):
                try:
                    b = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(d)).busId
                    b = (b.decode() if isinstance(b, bytes) else b).lower()
                    if len(b.split(":")[0]) == 8:
                        b = b[4:]
                    with open(f"/sys/bus/pci/devices/{b}/local_cpulist") as f:
                        if f.read().strip() == spec:
                            same.append(d)
                except Exception:
                    pass
            if device_index in same and len(same) > 1 and len(cpus) >= len(same):
                i, n = same.index(device_index), len(same)
                cpus = cpus[len(cpus) * i // n: len(cpus) * (i + 1) // n]
        os.sched_setaffinity(0, set(cpus))
        return cpus
    except Exception:
        return None


def partition_ranges(parts_per_layer: List[int], world: int) -> List[Tuple[int, int]]:
    """Split the flattened (layer, partition) list into `world` contiguous [begin, end) ranges balanced by
    multiply-accumulate work: a partition of layer l costs (P_l + 1) bins per P_l samples, i.e. the same per
    output sample for every layer, so the balance is by partition count.  Ranges may be empty when there are
    more ranks than partitions."""
    total = sum(parts_per_layer)
    out = []
    for r in range(world):
        out.append((total * r // world, total * (r + 1) // world))
    return out


def layer_slices(parts_per_layer: List[int], begin: int, end: int) -> List[Tuple[int, int]]:
    """Per-layer [q_begin, q_end) covered by the flattened range [begin, end) (what cpq_set_partition_range applies)."""
    out, base = [], 0
    for q in parts_per_layer:
        out.append((min(max(begin - base, 0), q), min(max(end - base, 0), q)))
        base += q
    return out


def reduce_partials(partial, op_group=None, owner_of_row=None):
    """Sum the ranks' partial wet signals in place.

    `partial` is a torch tensor [n_seq, T] (CUDA + NCCL on the box, CPU + gloo in tests).  With
    `owner_of_row` = None every rank gets the full sum (all-reduce).  Otherwise rows are reduced to their owning
    rank only (reduce per owner: the reduce-scatter-by-channel of SURVEY 8e) and other ranks' rows are left
    undefined."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(op_group) == 1:
        return partial
    if owner_of_row is None:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=op_group)
        return partial
    world = dist.get_world_size(op_group)
    for owner in range(world):
        rows = [i for i, o in enumerate(owner_of_row) if o == owner]
        if not rows:
            continue
        lo, hi = min(rows), max(rows) + 1
        if rows != list(range(lo, hi)):
            raise ValueError("rows of one owner must be contiguous")
        dist.reduce(partial[lo:hi], dst=owner, op=dist.ReduceOp.SUM, group=op_group)
    return partial


def process_partition_sharded_fused(engine, symm_tensor, peer_ptrs, T: int, rank: int, world: int, owned_streams, barrier,
                                    stages_after=None):
    """cfg 5 without a separate collective: every rank convolves its partition range into its own symmetric-memory buffer,
    then finishes the streams it owns while its EQ launch loads -- and sums, in rank order -- the tiles of ALL ranks'
    buffers over NVLink (cpq_set_partial_sources).  The reduce-scatter is the load stage of the consumer kernel.

    symm_tensor  this rank's [n_seq, stride] CUDA tensor allocated from torch symmetric memory, holding the full input
    peer_ptrs    device addresses of every rank's buffer in this process (rank order, own included)
    owned_streams (first, count) of the streams this rank finishes
    barrier      callable ordering the ranks' device work (e.g. the symmetric-memory handle's barrier)"""
    from . import capi
    lay = engine.layout()
    parts = [lay.layers[li].num_parts_ir for li in range(lay.num_layers)]
    b, e = partition_ranges(parts, world)[rank]
    engine.set_partition_range(b, e)
    stride = symm_tensor.stride(0)
    engine.process_device(symm_tensor.data_ptr(), stride, T, capi.STAGE_CONV)   # returns after the kernels have completed
    barrier()                                                                   # every rank's partial is in place
    engine.set_partition_range(0, -1)
    engine.set_partial_sources(list(peer_ptrs))
    engine.set_stream_window(*owned_streams)
    after = capi.STAGE_EQ | capi.STAGE_EPILOGUE if stages_after is None else stages_after
    engine.process_device(symm_tensor.data_ptr(), stride, T, after)
    engine.set_partial_sources([])
    engine.set_stream_window(0, -1)
    barrier()                                                                   # peers are done reading this rank's partial
    return symm_tensor


def process_partition_sharded(engine, io_tensor, T: int, rank: int, world: int, owner_of_row=None, stages_after=None):
    """cfg 5: convolve with this rank's partition range, sum the partials over ranks, then EQ + epilogue.

    `engine` is a ConvoPeqEngine whose impulses are the *full* IRs (every rank prepares the same engine);
    `io_tensor` is this rank's CUDA tensor [n_seq, stride] holding the full input (replicated)."""
    from . import capi
    parts = []
    lay = engine.layout()
    for li in range(lay.num_layers):
        parts.append(lay.layers[li].num_parts_ir)
    ranges = partition_ranges(parts, world)
    b, e = ranges[rank]
    engine.set_partition_range(b, e)
    stride = io_tensor.stride(0)
    engine.process_device(io_tensor.data_ptr(), stride, T, capi.STAGE_CONV)
    reduce_partials(io_tensor, owner_of_row=owner_of_row)
    if io_tensor.is_cuda:
        # the collective runs on torch's stream, the engine on its own: order them before the EQ stage reads the sum
        import torch
        torch.cuda.current_stream(io_tensor.device).synchronize()
    after = capi.STAGE_EQ | capi.STAGE_EPILOGUE if stages_after is None else stages_after
    if after:
        engine.process_device(io_tensor.data_ptr(), stride, T, after)
    engine.set_partition_range(0, -1)
    return io_tensor


def cfg5_layout(world: int):
    """BASELINE config 5 on `world` GPUs: the 8 channels are 4 stereo pairs that share one 2,097,152-tap IR pair.  Ranks form
    groups of two; a group owns 4 / (world / 2) pairs, and inside a group the two ranks split the flattened (layer, partition)
    list in half -- so the forward transforms of a channel run on two GPUs instead of on all of them, the multiply-accumulate
    is divided by the whole world, and each rank finishes (EQ, output stage) half of its group's channels after ONE pairwise
    reduce over NVLink.  Returns (groups, ranks_per_group, pairs_per_group)."""
    if world < 2 or world % 2 or 4 % (world // 2):
        raise ValueError("cfg5 sharding: world must be 2, 4 or 8")
    return world // 2, 2, 4 // (world // 2)


def bench_cfg5_sharded(local: int, rank: int, world: int, steps: int = 3):
    """Time cfg5 (192 kHz, 8 channels, 2M-tap IR, 10 s) partition-range sharded over `world` GPUs, with the one-GPU time of the
    same job (rank 0 alone) beside it.  Device-resident, CUDA events on the engine stream, max over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from . import capi
    from .engine import ConvoPeqEngine
    from tests import signals

    dev = torch.device("cuda", local)
    sr, block, T, ir_len = 192000.0, 512, 1920000 // 512 * 512, 2097152
    groups, rpg, pairs = cfg5_layout(world)
    grp, sub = rank // rpg, rank % rpg
    pgs = [dist.new_group(list(range(g * rpg, (g + 1) * rpg))) for g in range(groups)]
    irs = [signals.synth_ir(ir_len, 40 + ch) for ch in range(2)]
    spec = capi.default_filter_spec(sample_rate=sr)

    def build(n_pairs, first_pair):
        eng = ConvoPeqEngine(n_pairs, 2, sr, block, T, device=local, conv_boundary=capi.CONV_OUTER, shared_ir=True)
        for ch in range(2):
            eng.set_impulse(-1, ch, irs[ch], 1.0, spec)
        for s in range(n_pairs):
            eng.set_eq(s, signals.to_band(signals.band_params(seed=7 + first_pair + s)))
        eng.set_epilogue(1.0, 0)
        return eng

    def allmax(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    g = torch.Generator(device=dev)
    g.manual_seed(500)
    x_all = torch.randn(8, T, device=dev, dtype=torch.float64, generator=g) * 0.1     # the same on every rank
    eng = build(pairs, grp * pairs)
    x = x_all[2 * grp * pairs: 2 * (grp + 1) * pairs].contiguous()
    io = torch.empty_like(x)
    stream = torch.cuda.ExternalStream(eng.cuda_stream(), device=dev)
    lay = eng.layout()
    parts = [lay.layers[i].num_parts_ir for i in range(lay.num_layers)]
    b, e = partition_ranges(parts, rpg)[sub]
    # rows a rank finishes: half of the group's channels (whole streams when the group has two or more pairs)
    n_rows = 2 * pairs
    own_lo, own_hi = n_rows * sub // rpg, n_rows * (sub + 1) // rpg
    whole_streams = (own_lo % 2 == 0 and own_hi % 2 == 0)

    def step():
        io.copy_(x)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.set_partition_range(b, e)
        eng.process_device(io.data_ptr(), T, T, capi.STAGE_CONV)               # returns when the partial is complete
        dist.all_reduce(io, op=dist.ReduceOp.SUM, group=pgs[grp])             # pairwise sum over NVLink
        torch.cuda.current_stream(dev).synchronize()
        eng.set_partition_range(0, -1)
        if whole_streams:
            eng.set_stream_window(own_lo // 2, (own_hi - own_lo) // 2)
        eng.process_device(io.data_ptr(), T, T, capi.STAGE_EQ | capi.STAGE_EPILOGUE)
        eng.set_stream_window(0, -1)
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1)

    step()
    ms = [allmax(step()) for _ in range(steps)]
    sharded_ms = sum(ms) / len(ms)
    result_rows = io[own_lo:own_hi].clone() if whole_streams else io.clone()
    eng.close()
    # the same job on one GPU (rank 0; the other ranks wait)
    one_ms, err = None, None
    if rank == 0:
        full = build(4, 0)
        io1 = torch.empty_like(x_all)
        st1 = torch.cuda.ExternalStream(full.cuda_stream(), device=dev)

        def one():
            io1.copy_(x_all)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st1)
            full.process_device(io1.data_ptr(), T, T, capi.STAGE_ALL)
            e1.record(st1)
            e1.synchronize()
            return e0.elapsed_time(e1)

        one()
        o = [one() for _ in range(steps)]
        one_ms = sum(o) / len(o)
        lo = 2 * grp * pairs + own_lo
        ref = io1[lo: lo + result_rows.shape[0]] if whole_streams else io1[2 * grp * pairs: 2 * (grp + 1) * pairs]
        err = float((ref - result_rows).abs().max().item())
        full.close()
    dist.barrier()
    cs = 8.0 * T
    return {"workload": "cfg5: 192 kHz, 8 channels (4 stereo pairs sharing one 2,097,152-tap IR pair), block 512, 10 s, conv -> EQ -> makeup+headroom",
            "sharding": f"{groups} group(s) of {rpg} ranks; a group owns {pairs} pair(s) and splits the {sum(parts)} partitions "
                        f"({'+'.join(map(str, parts))}) in two; one pairwise NCCL all-reduce of the partial wet signal per group",
            "n_gpus": world, "device_ms": sharded_ms, "value": cs / (sharded_ms * 1e-3), "one_gpu_ms": one_ms,
            "speedup_vs_one_gpu": (one_ms / sharded_ms) if one_ms else None,
            "max_abs_diff_vs_one_gpu_rank0": err}
