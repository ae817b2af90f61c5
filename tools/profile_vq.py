"""Run the one-kernel VectorQuantizer a few times through the C ABI (for `ncu --set full -k regex:vq_fused`)."""
import ctypes
import sys

import torch

sys.path.insert(0, '.')
from codeformer_b200 import _lib  # noqa: E402

lib = _lib.load()
g = torch.Generator().manual_seed(0)
E = torch.randn(1024, 256, generator=g).cuda()
z = torch.randn(32, 256, 16, 16, generator=g).cuda()
prep = torch.empty(int(lib.cfb_vq_prepared_bytes(1024, 256)), dtype=torch.uint8, device='cuda')
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
_lib.check(lib.cfb_vq_prepare(_lib.ptr(E), 1024, 256, _lib.ptr(prep), prep.numel(), st))
zq = torch.empty_like(z)
idx = torch.empty((8192, 1), dtype=torch.int64, device='cuda')
stats = torch.empty(4, device='cuda')
ws = torch.empty(int(lib.cfb_vq_fast_workspace_bytes(32, 256, 256, 1024)), dtype=torch.uint8, device='cuda')
for _ in range(6):
    _lib.check(lib.cfb_vq_nearest_fast(_lib.ptr(z), _lib.ptr(E), _lib.ptr(prep), 32, 16, 16, 256, 1024, 0.25, _lib.ptr(zq), _lib.ptr(idx),
                                       _lib.ptr(stats), None, _lib.ptr(ws), ws.numel(), st))
torch.cuda.synchronize()
print('stats', stats.tolist())
