"""Would L2-resident sub-batches pay?  GroupNorm-affine + SiLU operand prep + 3x3 conv, chained x -> y -> x, per face,
at batch sizes whose working set does / does not fit the 126 MB L2; and the conv kernel alone for the same sizes."""
import ctypes
import json
import math
import sys

import torch

sys.path.insert(0, '.')
from codeformer_b200 import _lib  # noqa: E402

lib = _lib.load()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for (C, H) in [(128, 256), (64, 512), (256, 128)]:
    w = (torch.randn(C, C, 3, 3) / math.sqrt(9 * C)).cuda()
    b = torch.zeros(C, device='cuda')
    for N in [1, 2, 4, 8]:
        x = torch.randn(N, H, H, C, device='cuda')
        y = torch.empty_like(x)
        sc = torch.ones(N, C, device='cuda')
        sh = torch.zeros(N, C, device='cuda')
        wsb = lib.cfb_conv2d_workspace_bytes(N, H, H, C, C, 3, 0)
        ws = torch.empty(int(wsb), dtype=torch.uint8, device='cuda')

        def step(a, o):
            _lib.check(lib.cfb_conv2d_nhwc(_lib.ptr(a), _lib.ptr(w), _lib.ptr(b), _lib.ptr(o), N, H, H, C, C, 3, 0,
                                           _lib.ptr(sc), _lib.ptr(sh), 1, None, 0, 2, _lib.ptr(ws), wsb, st))
        reps = max(4, 32 // N)
        for _ in range(2):
            step(x, y); step(y, x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            step(x, y); step(y, x)
        e1.record()
        torch.cuda.synchronize()
        full = e0.elapsed_time(e1) / (2 * reps)
        ms = ctypes.c_float(0)
        _lib.check(lib.cfb_debug_time_conv(_lib.ptr(x), _lib.ptr(w), _lib.ptr(y), N, H, H, C, C, 3, 0, 2 * reps, _lib.ptr(ws), wsb, st, None, None, 0,
                                           ctypes.byref(ms)))
        print(json.dumps({'shape': f'{C}->{C}@{H}^2', 'batch': N, 'prep+conv_us_per_face': round(full / N * 1e3, 1),
                          'conv_only_us_per_face': round(ms.value / N * 1e3, 1),
                          'prep_us_per_face': round((full - ms.value) / N * 1e3, 1)}), flush=True)
