"""world_size-2 gloo tests of the N>1 host logic (no GPU): stream sharding needs no collective; partition-range
sharding sums rank partials (linearity of the convolver in the IR).  The arithmetic on each rank is done by the
CPU oracle here -- on the box the same plumbing drives the CUDA engine (tests/test_gpu_parity.py covers its
partition ranges on one GPU, bench.py / scripts/multi_gpu_check.py on several)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from convopeq_b200.dist import stream_range, partition_ranges, layer_slices, reduce_partials  # noqa: E402
from convopeq_b200.engine import plan_layout  # noqa: E402
from tests import signals  # noqa: E402


def test_stream_range_partitions_everything():
    for n in (1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            spans = [stream_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_partition_ranges_cover_the_layer_list():
    parts = [32, 64, 56]        # cfg5: L0 32 x 512, L1 64 x 4096, L2 56 x 32768
    for world in (1, 2, 4, 8, 200):
        rs = partition_ranges(parts, world)
        assert rs[0][0] == 0 and rs[-1][1] == sum(parts)
        covered = [0, 0, 0]
        for b, e in rs:
            for li, (qb, qe) in enumerate(layer_slices(parts, b, e)):
                covered[li] += qe - qb
        assert covered == parts


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.bindings import Oracle
        orc = Oracle()
        # ---- partition-range sharding: rank keeps only the IR samples of its partitions, partials are all-reduced ----
        ir_len, block, T = 40000, 256, 16384
        ir = signals.synth_ir(ir_len, 21)
        x = signals.noise(T, 22)
        lay, _ = plan_layout(ir_len, block, None, T // block)
        parts = [lay.layers[i].num_parts_ir for i in range(lay.num_layers)]
        b, e = partition_ranges(parts, world)[rank]
        masked = np.zeros_like(ir)
        for li, (qb, qe) in enumerate(layer_slices(parts, b, e)):
            L = lay.layers[li]
            lo = L.ir_offset + qb * L.part_size
            hi = min(L.ir_offset + qe * L.part_size, L.ir_offset + L.ir_len)
            if hi > lo:
                masked[lo:hi] = ir[lo:hi]
        masked[-1] += 0.0   # same length -> same layer plan on every rank
        part, lay_r = orc.nuc_run(masked, x, block)
        t = torch.from_numpy(part.copy())[None, :]
        reduce_partials(t)
        full, _ = orc.nuc_run(ir, x, block)
        err_part = float(np.abs(t[0].numpy() - full).max())
        # ---- reduce-to-owner variant ----
        t2 = torch.from_numpy(np.stack([part, part]))
        reduce_partials(t2, owner_of_row=[0, 1])
        err_owner = float(np.abs(t2[rank].numpy() - full).max())
        # ---- stream sharding: no collective on the data path, results gathered only for the check ----
        n_streams = 5
        sb, se = stream_range(n_streams, rank, world)
        mine = {}
        for s in range(sb, se):
            y, _ = orc.nuc_run(signals.synth_ir(3000, 100 + s), signals.noise(4096, 200 + s), block)
            mine[s] = float(np.abs(y).sum())
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        q.put((rank, err_part, err_owner, gathered))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err_part, err_owner, gathered in res:
        assert err_part <= 1e-12, err_part
        assert err_owner <= 1e-12, err_owner
        owned = sorted(k for d in gathered for k in d)
        assert owned == [0, 1, 2, 3, 4]          # every stream processed exactly once across ranks
