"""Which smem rows/chunks does a row-shifted SWIZZLE_128B K-major UMMA descriptor read?  (run on the B200 box)"""
import ctypes
import sys

import torch

sys.path.insert(0, '.')
from codeformer_b200 import _lib  # noqa: E402

lib = _lib.load()
R = 256
r = torch.arange(R).view(R, 1).float()
c = (torch.arange(64) // 8).view(1, 64).float()
A = (r + 256 * c).half().cuda()                     # value = row + 256*chunk (exact in fp16)
B = torch.eye(64).half().cuda()
cfgs = []
for shift in (0, 1, 2, 3, 7, 8, 9):
    for boff in sorted({0, shift % 8}):
        cfgs.append((shift, boff, 1024))
for shift in (0, 1, 2):                              # 8-row groups 16 rows apart (halo-patch layout)
    for boff in sorted({0, shift % 8}):
        cfgs.append((shift, boff, 2048))
cfg = torch.tensor(cfgs, dtype=torch.int32).cuda()
out = torch.zeros(len(cfgs), 128, 64, device='cuda')
_lib.check(lib.cfb_debug_umma_probe(_lib.ptr(A), R, _lib.ptr(B), _lib.ptr(cfg), len(cfgs), _lib.ptr(out),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
torch.cuda.synchronize()
out = out.cpu()
for i, (shift, boff, sbo) in enumerate(cfgs):
    m = torch.arange(128)
    exp_row = shift + (m // 8) * (sbo // 128) + m % 8
    exp = exp_row.view(128, 1).float() + 256 * (torch.arange(64) // 8).view(1, 64).float()
    ok = bool(torch.equal(out[i], exp))
    rows = (out[i] % 256)
    chunks = (out[i] // 256)
    msg = f'shift={shift} base_offset={boff} sbo={sbo}: {"OK" if ok else "MISMATCH"}'
    if not ok:
        msg += f' | rows m=0..9 (col0): {rows[:10, 0].int().tolist()} | chunk ids of row 0: {chunks[0, ::8].int().tolist()}' \
               f' row 1: {chunks[1, ::8].int().tolist()}'
    print(msg)
