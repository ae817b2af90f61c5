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
