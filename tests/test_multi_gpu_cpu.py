"""N>1 host logic on CPU: gloo backend, world_size 2 (the NCCL path in bench.py runs the same functions).
Checks the frame dealing, the owner rotation and that after the exchange every rank holds the owner's
reference plane bit-exactly (results then cannot depend on N: each frame's arithmetic stays on its rank)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from x265_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    seen = []
    for step in range(4):
        owner = shard.ref_owner(step, world)
        plane = torch.full((64 * 80,), 255, dtype=torch.uint8)
        truth = torch.from_numpy(np.random.default_rng(1000 + step).integers(0, 256, 64 * 80).astype(np.uint8))
        if rank == owner:
            plane.copy_(truth)                      # only the owner has the reconstructed pixels
        shard.exchange_ref(dist, plane, step, world)
        ok = ok and bool(torch.equal(plane, truth))
        # incoming-reference variant: non-owners receive into `recv`, their own plane stays as it was
        mine = torch.full((64 * 80,), rank + 1, dtype=torch.uint8)
        recv = torch.zeros(64 * 80, dtype=torch.uint8)
        got = shard.exchange_ref(dist, truth.clone() if rank == owner else mine, step, world, recv=recv)
        if rank == owner:
            ok = ok and bool(torch.equal(got, truth)) and int(recv.sum()) == 0
        else:
            ok = ok and bool(torch.equal(recv, truth)) and bool(torch.equal(mine, torch.full_like(mine, rank + 1))) and got is recv
        seen.append(shard.frame_of(step, rank, world))
    gathered = [None] * world
    dist.all_gather_object(gathered, seen)
    if rank == 0:
        allf = sorted(f for g in gathered for f in g)
        ok = ok and allf == list(range(4 * world))   # every frame exactly once
    t = torch.tensor([1 if ok else 0]); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ret[rank] = int(t.item())
    dist.destroy_process_group()


def test_sharding_and_exchange_gloo():
    world = 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] == 1 for r in range(world))


def test_frame_dealing_is_a_partition():
    for world in (1, 2, 4, 8):
        frames = sorted(shard.frame_of(s, r, world) for s in range(5) for r in range(world))
        assert frames == list(range(5 * world))
        assert [shard.ref_owner(s, world) for s in range(world)] == list(range(world))


def _row_worker(rank, world, port, ret, mode):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, stride, my, nrows = 136, 448, 80, 3          # a 200x136 picture: 3 CTU rows, the last one 8 pixels high
    truth = torch.from_numpy(np.random.default_rng(7).integers(0, 256, stride * (H + 2 * my)).astype(np.uint8))
    plane = torch.zeros_like(truth)
    for r0, r1 in shard.row_blocks(nrows, rank, world, mode):      # a rank holds only the rows it reconstructed
        b0, b1 = shard.band_slice(r0, r1, H, stride, my)
        plane[b0:b1] = truth[b0:b1]
    shard.exchange_rows(dist, plane, nrows, world, H, stride, my, mode=mode)
    lo, hi = my * stride, (my + H) * stride
    ok = bool(torch.equal(plane[lo:hi], truth[lo:hi])) and int(plane[:lo].sum()) == 0 and int(plane[hi:].sum()) == 0
    t = torch.tensor([1 if ok else 0]); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ret[rank] = int(t.item())
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["block", "cyclic"])
def test_row_shard_exchange_gloo(mode):
    world = 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_row_worker, args=(world, _free_port(), ret, mode), nprocs=world, join=True)
    assert all(ret[r] == 1 for r in range(world))


@pytest.mark.parametrize("mode", ["block", "cyclic"])
def test_row_blocks_are_a_partition(mode):
    for nrows in (1, 3, 34, 68):
        for world in (1, 2, 4, 8):
            rows = sorted(r for g in range(world) for r0, r1 in shard.row_blocks(nrows, g, world, mode) for r in range(r0, r1))
            assert rows == list(range(nrows))
            assert all(shard.row_owner(r, nrows, world, mode) in range(world) for r in range(nrows))
            sizes = [sum(r1 - r0 for r0, r1 in shard.row_blocks(nrows, g, world, mode)) for g in range(world)]
            assert max(sizes) - min(sizes) <= 1


# ---- lookahead frames sharded per rank (BASELINE configs[3]; x265_b200/lookahead.py) ----------------------------------
def _la_worker(rank, world, port, ret):
    """Every rank builds the lowres planes of ITS frames only (owner(b) = b % world), publishes them with one broadcast per
    frame, then estimates the triples whose b it owns (CPU oracle arithmetic here; the GPU path runs the same host logic
    over NCCL).  The gathered costs must equal a single-process run: sharding changes placement, never arithmetic."""
    import ctypes as C
    from common import load_oracle
    from frame_helpers import gen_luma
    from lookahead_helpers import OracleLookahead, make_lowres
    from x265_b200.lookahead import owner, window_triples, conflict_free_batches
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O = load_oracle(8)
    W, H, n = 208, 144, 6
    frames = [gen_luma(W, H, i, s1=17.0, s2=11.0) for i in range(n)]
    blank = [np.zeros_like(f) for f in frames]
    # a rank only has the source pixels of its own frames
    la = OracleLookahead(O, [frames[i] if owner(i, world) == rank else blank[i] for i in range(n)], 8)
    for i in range(n):
        for k in range(4):
            t = torch.from_numpy(la.fr[i]["planes"][k])
            dist.broadcast(t, src=owner(i, world))             # one frame's planes from its owner
    triples = window_triples(n, 2)
    mine = [t for t in triples if owner(t[2], world) == rank]
    # intra costs of the frames this rank estimates are its own (it owns b); p0 / p1 only contribute planes
    costs = {t: la.cost(*t) for t in mine}
    gathered = [None] * world
    dist.all_gather_object(gathered, costs)
    ok = 1
    if rank == 0:
        full = OracleLookahead(O, frames, 8)
        want = {t: full.cost(*t) for t in triples}
        got = {}
        for g in gathered:
            got.update(g)
        ok = int(got == want and len(got) == len(triples))
        # launches never put two searches of one motion field together
        for b in conflict_free_batches(triples):
            keys = [(t[2], 0, t[2] - t[0]) for t in b] + [(t[2], 1, t[1] - t[2]) for t in b if t[1] > t[2]]
            ok = ok and int(len(keys) == len(set(keys)))
    t = torch.tensor([ok]); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ret[rank] = int(t.item())
    dist.destroy_process_group()


def test_lookahead_frame_shards_gloo():
    world = 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_la_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] == 1 for r in range(world))


def test_lookahead_owner_is_a_partition():
    from x265_b200.lookahead import owner, window_triples
    for world in (1, 2, 4, 8):
        tr = window_triples(20, 4)
        per = [[t for t in tr if owner(t[2], world) == g] for g in range(world)]
        assert sorted(t for p in per for t in p) == sorted(tr)
        assert max(len(p) for p in per) - min(len(p) for p in per) <= len(tr) // world // 2 + 16
