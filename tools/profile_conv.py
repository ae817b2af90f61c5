"""Run one conv through the C ABI a few times (for `ncu --set full -k regex:conv_tc`)."""
import argparse
import ctypes
import math
import sys

import torch

sys.path.insert(0, '.')
from codeformer_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=8)
ap.add_argument('--cin', type=int, default=128)
ap.add_argument('--cout', type=int, default=128)
ap.add_argument('--h', type=int, default=256)
ap.add_argument('--k', type=int, default=3)
ap.add_argument('--mode', type=int, default=0)
ap.add_argument('--reps', type=int, default=4)
ap.add_argument('--residual', type=int, default=0)
ap.add_argument('--xf', type=int, default=0, help='GroupNorm-affine + SiLU input: the fused operand transform variant')
a = ap.parse_args()
lib = _lib.load()
N, H, C1, C2 = a.n, a.h, a.cin, a.cout
x = torch.randn(N, H, H, C1, device='cuda')
w = (torch.randn(C2, C1, a.k, a.k) / math.sqrt(a.k * a.k * C1)).cuda()
b = torch.zeros(C2, device='cuda')
Ho = H // 2 if a.mode == 1 else (H * 2 if a.mode == 2 else H)
out = torch.empty(N, Ho, Ho, C2, device='cuda')
wsb = lib.cfb_conv2d_workspace_bytes(N, H, H, C1, C2, a.k, a.mode)
ws = torch.empty(int(wsb), dtype=torch.uint8, device='cuda')
res = torch.randn(N, Ho, Ho, C2, device='cuda') if a.residual else None
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
sc = (1 + 0.1 * torch.randn(N, C1, device='cuda')) if a.xf else None
sh = (0.1 * torch.randn(N, C1, device='cuda')) if a.xf else None
for _ in range(a.reps):
    _lib.check(lib.cfb_conv2d_nhwc(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(out), N, H, H, C1, C2, a.k, a.mode,
                                   _lib.ptr(sc), _lib.ptr(sh), 1 if a.xf else 0, _lib.ptr(res), 0, 2, _lib.ptr(ws), wsb, st))
torch.cuda.synchronize()
print('done')
