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
