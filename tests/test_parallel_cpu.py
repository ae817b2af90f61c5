"""Host logic of the multi-GPU path on CPU: world_size-2 gloo ranks, shard -> compute -> all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from codeformer_b200.parallel import StreamedGather, gather_faces, pipelined_forward_gather, shard_bounds, sharded_forward


def test_shard_bounds_cover_and_balance():
    for B in (0, 1, 2, 5, 32, 255, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


class _FakeNet:
    """Stands in for the CUDA module: any per-face deterministic map exercises the sharding logic."""

    def __call__(self, x, w=0.5):
        return (x * w + x.flatten(1).sum(1).view(-1, 1, 1, 1),)


def _worker(rank, world, port, batch, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(batch, 3, 8, 8, generator=g)
        full = _FakeNet()(x, w=0.5)[0]
        out = sharded_forward(_FakeNet(), x, w=0.5)
        ok = torch.equal(out, full)                       # gathered result is bit-identical to the unsharded run
        lo, hi = shard_bounds(batch, rank, world)
        ok = ok and torch.equal(gather_faces(full[lo:hi], batch), full)
        # pipelined variant (equal shards per rank): sub-batch k's gather overlaps sub-batch k+1's compute; same bits
        per = 3
        xg = torch.randn(world * per, 3, 8, 8, generator=g)
        fullg = _FakeNet()(xg, w=0.5)[0]
        for chunks in (1, 2, 3):
            got, mine = pipelined_forward_gather(_FakeNet(), xg[rank * per:(rank + 1) * per], chunks=chunks, w=0.5)
            ok = ok and torch.equal(got, fullg) and torch.equal(mine, fullg[rank * per:(rank + 1) * per])
        # stream of batches: gather i overlaps forward i+1, results arrive one submit later and equal the blocking gather
        sg, got = StreamedGather(), []
        for i in range(3):
            xi = xg + i
            prev = sg.submit(_FakeNet()(xi[rank * per:(rank + 1) * per], w=0.5)[0])
            if prev is not None:
                got.append(prev.clone())
        got.append(sg.flush().clone())
        ok = ok and len(got) == 3 and all(torch.equal(got[i], _FakeNet()(xg + i, w=0.5)[0]) for i in range(3))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('batch', [4, 5, 1])
def test_two_rank_gloo_gather_is_bit_identical(batch):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, batch, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_restore_faces_chunk_plan():
    """Host logic of the batched caller loop: chunks cover every face once, in order, within max_batch."""
    from codeformer_b200.arch import restore_chunks
    assert restore_chunks(0, 32) == []
    assert restore_chunks(7, 3) == [(0, 3), (3, 6), (6, 7)]
    assert restore_chunks(32, 32) == [(0, 16), (16, 32)]
    assert restore_chunks(33, 32) == [(0, 16), (16, 32), (32, 33)]
    for n in (1, 5, 15, 16, 17, 31, 64, 100):
        for mb in (1, 4, 16, 32):
            b = restore_chunks(n, mb)
            assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b, b[1:]))
            assert all(0 < hi - lo <= mb for lo, hi in b)
