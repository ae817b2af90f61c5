"""CTA-pair MMA probe (tcgen05.mma.cta_group::2, M=256 over the two SMs of a TPC): operand-half -> accumulator-column
mapping and sustained rate vs N (run on the B200 box).  A = rank+1, B half of rank r = r+1, K = 64*reps:
expected D[rows of rank r][col] = K * (r+1) * (1 if the column comes from rank 0's B half else 2)."""
import ctypes
import sys

import torch

sys.path.insert(0, '.')
from codeformer_b200 import _lib  # noqa: E402

lib = _lib.load()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for N in (64, 128, 256):
    vals = torch.zeros(2 * 2 * 4, dtype=torch.float32, device='cuda')
    info = torch.zeros(2 * 2, dtype=torch.int64, device='cuda')
    _lib.check(lib.cfb_debug_umma_pair(N, 2, _lib.ptr(vals), _lib.ptr(info), 2, st))
    torch.cuda.synchronize()
    v = vals.view(2, 2, 4).cpu()
    print(f'N={N}: K=128; rank0 lanes0/96 cols[0,31,N-32,N-1] = {v[0].tolist()}  rank1 = {v[1].tolist()}  tmem bases = '
          f'{[hex(int(x)) for x in info.view(2, 2)[:, 1].tolist()]}', flush=True)
reps = 2000
for ctas in (2, 148):
    for N in (64, 128, 256):
        vals = torch.zeros(ctas * 2 * 4, dtype=torch.float32, device='cuda')
        info = torch.zeros(ctas * 2, dtype=torch.int64, device='cuda')
        for _ in range(2):
            _lib.check(lib.cfb_debug_umma_pair(N, reps, _lib.ptr(vals), _lib.ptr(info), ctas, st))
        torch.cuda.synchronize()
        cyc = info.view(ctas, 2)[::2, 0].float().mean().item() / (reps * 4)
        ideal = 128 * N / 256           # per SM: 128 x N x 16 MACs at 8192 MACs/clk/SM
        print(f'ctas={ctas:3d} N={N:3d}: {cyc:7.1f} cycles per M=256 MMA (ideal {ideal:.0f}) -> {100 * ideal / cyc:5.1f}% of peak', flush=True)
