"""CPU, world_size 2 over gloo: the partition and exchange plumbing of the multi-GPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from particlesim_b200.parallel import all_gather_slices, shard_range, shard_width


def test_shard_ranges_cover_everything():
    for n in (0, 1, 2, 7, 100, 16_000_001):
        for world in (1, 2, 3, 8):
            w = shard_width(n, world)
            assert w * world >= n
            covered = 0
            for r in range(world):
                f, c = shard_range(n, world, r)
                assert f == min(r * w, n) and 0 <= c <= w and f + c <= n
                covered += c
            assert covered == n


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = shard_width(n, world)
        # every rank starts from the same replicated state and updates only its own slice
        full = torch.arange(world * w * 4, dtype=torch.float32).reshape(world * w, 4)
        f, c = shard_range(n, world, rank)
        full[f:f + c] += 1000.0 * (rank + 1)
        all_gather_slices(full, w, rank, world, dist)
        expect = torch.arange(world * w * 4, dtype=torch.float32).reshape(world * w, 4)
        for r in range(world):
            ff, cc = shard_range(n, world, r)
            expect[ff:ff + cc] += 1000.0 * (r + 1)
        ok = bool(torch.equal(full[:n], expect[:n]))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 11])
def test_all_gather_slices_gloo_world2(n):
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _worker_v(rank, world, port, q):
    """the exchanges of the sharded build (DistComm): uneven all-gather and the sum all-reduce whose slots
    have exactly one non-zero writer"""
    from particlesim_b200.parallel import DistComm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for padded, offsets in [(p, o) for p in (False, True) for o in ([0, 3, 10], [0, 0, 10], [0, 10, 10], [0, 7, 7])]:
            comm = DistComm(dist, rank, padded=padded)
            full = torch.full((10, 4), -1, dtype=torch.int32)
            a, b = offsets[rank], offsets[rank + 1]
            full[a:b] = 100 * (rank + 1) + torch.arange(b - a, dtype=torch.int32)[:, None]
            comm.all_gatherv([full], offsets)
            expect = torch.full((10, 4), -1, dtype=torch.int32)
            for r in range(world):
                aa, bb = offsets[r], offsets[r + 1]
                expect[aa:bb] = 100 * (r + 1) + torch.arange(bb - aa, dtype=torch.int32)[:, None]
            ok &= bool(torch.equal(full, expect))
        comm = DistComm(dist, rank)
        words = torch.zeros(64, dtype=torch.int64)
        words[rank::world] = torch.arange(64, dtype=torch.int64)[rank::world] * 0x0123456789 + 1
        comm.all_reduce([words])
        ok &= bool(torch.equal(words, torch.arange(64, dtype=torch.int64) * 0x0123456789 + 1))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_sharded_build_exchanges_gloo_world2():
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_v, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_loopback_comm_matches_the_collectives():
    from particlesim_b200.parallel import LoopbackComm
    comm = LoopbackComm()
    offsets = [0, 2, 2, 9]
    ts = []
    for r in range(3):
        t = torch.full((9,), -1, dtype=torch.int32)
        t[offsets[r]:offsets[r + 1]] = r
        ts.append(t)
    comm.all_gatherv(ts, offsets)
    for t in ts:
        assert t.tolist() == [0, 0, 2, 2, 2, 2, 2, 2, 2]
    ws = [torch.tensor([1, 0, 0], dtype=torch.int64), torch.tensor([0, 5, 0], dtype=torch.int64)]
    comm.all_reduce(ws)
    assert ws[0].tolist() == ws[1].tolist() == [1, 5, 0]
