"""GPU probe of the tcgen05 conv engine: parity vs torch CPU fp32 and vs the fp32 CUDA-core engine, plus timings.
Run on the B200 box:  python tools/tc_probe.py   (each case is isolated by try/except; a trap poisons the context)."""
import math
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
from tests import gpu_util as G  # noqa: E402
from codeformer_b200 import _lib  # noqa: E402

torch.set_grad_enabled(False)


def rnd(*s, seed=0, scale=1.0):
    return torch.randn(*s, generator=torch.Generator().manual_seed(seed)) * scale


CASES = [  # N, Cin, Cout, H, k, mode
    (1, 64, 64, 16, 1, 0), (1, 64, 64, 16, 3, 0), (1, 64, 128, 16, 3, 0), (2, 128, 128, 32, 3, 0), (1, 128, 64, 128, 3, 0),
    (1, 256, 256, 64, 3, 0), (1, 512, 512, 16, 3, 0), (1, 512, 1536, 16, 1, 0), (1, 128, 128, 16, 3, 2), (1, 64, 64, 256, 3, 0),
    (1, 64, 64, 32, 3, 1), (2, 128, 128, 256, 3, 1), (1, 256, 256, 64, 3, 1), (1, 512, 256, 16, 3, 0),
]


def main():
    import os
    print(torch.cuda.get_device_name(0), 'CFB_TC_CHUNK=', os.environ.get('CFB_TC_CHUNK'), flush=True)
    for (N, Cin, Cout, H, k, mode) in CASES:
        x = rnd(N, Cin, H, H, seed=1)
        w = rnd(Cout, Cin, k, k, seed=2, scale=1 / math.sqrt(Cin * k * k))
        b = rnd(Cout, seed=3, scale=0.1)
        if mode == 1:
            ref = F.conv2d(F.pad(x, (0, 1, 0, 1)).double(), w.double(), b.double(), stride=2).float()
        else:
            xin = F.interpolate(x, scale_factor=2.0, mode='nearest') if mode == 2 else x
            ref = F.conv2d(xin.double(), w.double(), b.double(), padding=k // 2).float()
        try:
            t = time.time()
            o2 = G.conv2d(x, w, b, mode=mode, engine=2).cpu()
            o1 = G.conv2d(x, w, b, mode=mode, engine=1).cpu()
            e2, e1 = float((o2 - ref).abs().max()), float((o1 - ref).abs().max())
            d = (o2 - ref).abs()
            bad = int((d > 1e-3).sum())
            print(f'case N{N} {Cin}->{Cout} H{H} k{k} m{mode}: tc err {e2:.3e} f32 err {e1:.3e} |ref|max {float(ref.abs().max()):.2f} '
                  f'bad>{1e-3}: {bad}/{d.numel()}  ({time.time() - t:.1f}s)', flush=True)
            if bad:
                idx = (d > 1e-3).nonzero()[:6].tolist()
                print('   first bad (n,c,y,x):', idx, 'got', [float(o2[tuple(i)]) for i in idx], 'want', [float(ref[tuple(i)]) for i in idx], flush=True)
        except Exception as e:  # noqa: BLE001
            print(f'case N{N} {Cin}->{Cout} H{H} k{k} m{mode}: EXCEPTION {e}', flush=True)
            break
    # timing of the dominant shapes
    lib = _lib.load()
    import ctypes
    for (N, C1, C2, H) in [(8, 128, 128, 256), (8, 64, 64, 512), (8, 256, 256, 64), (8, 512, 512, 16)]:
        for engine in (2, 1):
            try:
                x = torch.randn(N, H, H, C1, device='cuda')
                w = (torch.randn(C2, C1, 3, 3) / math.sqrt(9 * C1)).cuda()
                b = torch.zeros(C2, device='cuda')
                out = torch.empty(N, H, H, C2, device='cuda')
                wsb = lib.cfb_conv2d_workspace_bytes(N, H, H, C1, C2, 3, 0)
                ws = torch.empty(int(wsb), dtype=torch.uint8, device='cuda')
                st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

                def run():
                    _lib.check(lib.cfb_conv2d_nhwc(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(out), N, H, H, C1, C2, 3, 0,
                                                   None, None, 0, None, 0, engine, _lib.ptr(ws), wsb, st))
                for _ in range(2):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    run()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                fl = 2.0 * N * H * H * C1 * C2 * 9
                print(f'time engine{engine} N{N} {C1}->{C2}@{H}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (incl. prep+split)', flush=True)
            except Exception as e:  # noqa: BLE001
                print(f'time engine{engine} N{N} {C1}->{C2}@{H}: EXCEPTION {e}', flush=True)
                return


if __name__ == '__main__':
    main()
