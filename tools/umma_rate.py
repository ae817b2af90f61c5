"""Sustained tcgen05.mma rate vs N and accumulator rotation (run on the B200 box)."""
import ctypes
import sys

import torch

sys.path.insert(0, '.')
from codeformer_b200 import _lib  # noqa: E402

lib = _lib.load()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
reps = 400
for ctas in (1, 148):
    for N in (64, 128, 256):
        for nacc in (1, 2, 4):
            if nacc * N > 512:
                continue
            out = torch.zeros(ctas, dtype=torch.int64, device='cuda')
            for _ in range(2):
                _lib.check(lib.cfb_debug_umma_rate(N, nacc, reps, _lib.ptr(out), ctas, st))
            torch.cuda.synchronize()
            cyc = out.float().mean().item() / (reps * 12)
            ideal = 128 * N / 256
            print(f'ctas={ctas:3d} N={N:3d} nacc={nacc}: {cyc:7.1f} cycles/MMA (ideal {ideal:.0f}) -> {100 * ideal / cyc:5.1f}% of peak')
