"""-m gpu, needs >= 2 GPUs: the sharded run (one process per GPU, NCCL all-gather of `out`) must be bit-identical to
the single-GPU run on the same faces (SURVEY.md §4 iv, §8e)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import codeformer_b200 as cb
        from codeformer_b200 import spec as S
        from codeformer_b200.parallel import sharded_forward
        from tests.util import faces_input
        torch.set_grad_enabled(False)
        net = cb.CodeFormer().to(f'cuda:{rank}').eval()
        net.load_state_dict(S.random_state_dict(S.codeformer_spec(), 1))
        x = faces_input(slice(0, 4)).to(f'cuda:{rank}')
        full = sharded_forward(net, x, w=0.5, adain=True)
        single = net(x, w=0.5, adain=True)[0]
        ret[rank] = bool(torch.equal(full, single))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_gpu_gather_bit_identical():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def _worker32(rank, world, port, ret):
    """BASELINE configs[4] shape per rank (32 faces / GPU) through the pipelined front-end: the gathered tensor must hold every
    rank's result bit for bit (checked on every rank against its own shard and against a 1-rank recomputation of a sample)."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import codeformer_b200 as cb
        from codeformer_b200 import spec as S
        from codeformer_b200.parallel import pipelined_forward_gather
        torch.set_grad_enabled(False)
        net = cb.CodeFormer().to(f'cuda:{rank}').eval()
        net.load_state_dict(S.random_state_dict(S.codeformer_spec(), 1))
        per = 32
        g = torch.Generator().manual_seed(5)
        x_all = torch.randn(world * per, 3, 512, 512, generator=g).clamp_(-1, 1)
        x = x_all[rank * per:(rank + 1) * per].to(f'cuda:{rank}')
        full, mine = pipelined_forward_gather(net, x, chunks=2, w=0.5, adain=True)
        torch.cuda.synchronize()
        ok = torch.equal(full[rank * per:(rank + 1) * per], mine)
        other = (rank + 1) % world                                # recompute 4 faces of the neighbour's shard locally
        chk = net(x_all[other * per:other * per + 4].to(f'cuda:{rank}'), w=0.5, adain=True)[0]
        ok = ok and torch.equal(full[other * per:other * per + 4], chk)
        one = net(x, w=0.5, adain=True)[0]                        # one 32-face forward == two pipelined 16-face halves
        ok = ok and torch.equal(one, mine)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_gpu_pipelined_gather_32_faces_per_rank():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker32, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
