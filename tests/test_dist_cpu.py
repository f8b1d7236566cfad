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


def test_cfg5_hybrid_layout_covers_pairs_and_partitions():
    """dist.cfg5_layout: ranks in groups of two, a group owns 4 / (world / 2) stereo pairs, the two ranks of a group split
    the partition list in half -- every (pair, partition) is computed exactly once."""
    from convopeq_b200.dist import cfg5_layout
    parts = [32, 64, 56]
    for world in (2, 4, 8):
        groups, rpg, pairs = cfg5_layout(world)
        assert groups * rpg == world and groups * pairs == 4 and rpg == 2
        seen = {}
        for rank in range(world):
            grp, sub = rank // rpg, rank % rpg
            b, e = partition_ranges(parts, rpg)[sub]
            for pair in range(grp * pairs, (grp + 1) * pairs):
                for q in range(b, e):
                    assert (pair, q) not in seen
                    seen[(pair, q)] = rank
        assert len(seen) == 4 * sum(parts)
    for bad in (1, 3, 6, 16):
        with pytest.raises(ValueError):
            cfg5_layout(bad)


def _group_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from convopeq_b200.dist import cfg5_layout
        from oracle.bindings import Oracle
        orc = Oracle()
        groups, rpg, pairs = cfg5_layout(world)
        grp, sub = rank // rpg, rank % rpg
        pgs = [dist.new_group(list(range(g * rpg, (g + 1) * rpg))) for g in range(groups)]
        ir_len, block, T = 30000, 256, 8192
        ir = signals.synth_ir(ir_len, 31)
        lay, _ = plan_layout(ir_len, block, None, T // block)
        parts = [lay.layers[i].num_parts_ir for i in range(lay.num_layers)]
        b, e = partition_ranges(parts, rpg)[sub]
        masked = np.zeros_like(ir)
        for li, (qb, qe) in enumerate(layer_slices(parts, b, e)):
            L = lay.layers[li]
            lo, hi = L.ir_offset + qb * L.part_size, min(L.ir_offset + qe * L.part_size, L.ir_offset + L.ir_len)
            if hi > lo:
                masked[lo:hi] = ir[lo:hi]
        # the group's channels: pair p = channels 2p, 2p + 1 of the 8-channel stream
        chans = range(2 * grp * pairs, 2 * (grp + 1) * pairs)
        x = np.stack([signals.noise(T, 100 + c) for c in chans])
        # the restatement has no partition range: a rank's partial is the full plan run on the IR with the other partitions
        # zeroed (same layer plan, the convolver is linear in the IR)
        part = np.stack([orc.nuc_run(masked, x[i], block)[0] for i in range(x.shape[0])])
        t = torch.from_numpy(part)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=pgs[grp])
        want = np.stack([orc.nuc_run(ir, x[i], block)[0] for i in range(x.shape[0])])
        q.put((rank, float(np.abs(t.numpy() - want).max())))
    finally:
        dist.destroy_process_group()


def test_cfg5_hybrid_group_reduce_two_ranks_gloo():
    """world_size 2 = one group: each rank convolves its half of the partitions, the pairwise all-reduce gives the full sum."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_group_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(err <= 1e-12 for _, err in res), res
