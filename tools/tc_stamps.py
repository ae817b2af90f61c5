"""Where does a small tcgen05 conv launch spend its time?  For each shape: back-to-back launch time (CUDA events, through
cfb_debug_time_conv) and the phase stamps of CTA 0 of the last launch (cfb_debug_set_stamps, SM cycles).

The stamps are a BUILD option of the library (they cost 6-13 % of every conv kernel even when unused, see csrc/conv_tc.cu):
    tools/build_ab.sh stamps -DCFB_TC_STAMPS=1 && CFB_LIB=$PWD/codeformer_b200/ab/lib_stamps.so python tools/tc_stamps.py
With the production build the launch times are printed and every stamp reads zero."""
import ctypes
import sys

import torch

sys.path.insert(0, '.')
from codeformer_b200 import _lib  # noqa: E402

lib = _lib.load()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
NAMES = {1: 'setup done', 2: 'first TMA issued', 10: 'xf: first raw patch requested', 11: 'xf: raw patch landed', 12: 'xf: first patch transformed',
         3: 'first MMA issued', 4: 'last MMA of tile 0 issued', 5: 'epilogue: first accumulator', 6: 'epilogue: last accumulator',
         7: 'epilogue: K loop of tile 0 done', 16: 'epilogue: tile 0 stored', 17: 'epilogue: K loop of tile 1 done', 18: 'epilogue: tile 1 stored',
         19: 'xf: second tile patch landed', 20: 'xf: second tile patch transformed', 13: 'epilogue: all tiles stored', 8: 'tear-down sync', 9: 'cluster sync'}


def run(n, h, w, cin, cout, k, xf, reps=200):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, h, w, cin, generator=g).cuda()
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    out = torch.empty(n, h, w, cout, device='cuda')
    ws = torch.empty(int(lib.cfb_conv2d_workspace_bytes(n, h, w, cin, cout, k, 0)), dtype=torch.uint8, device='cuda')
    sc = torch.rand(n, cin, generator=g).cuda() + 0.5
    sh = torch.randn(n, cin, generator=g).cuda() * 0.1
    stamps = torch.zeros(32, dtype=torch.int64, device='cuda')
    ms = ctypes.c_float(0)
    args = (_lib.ptr(x), _lib.ptr(wt), _lib.ptr(out), n, h, w, cin, cout, k, 0, reps, _lib.ptr(ws), ws.numel(), st,
            _lib.ptr(sc) if xf else None, _lib.ptr(sh) if xf else None, 1 if xf else 0, ctypes.byref(ms))
    _lib.check(lib.cfb_debug_time_conv(*args), 'time_conv')
    base = ms.value * 1e3
    _lib.check(lib.cfb_debug_set_stamps(_lib.ptr(stamps)), 'set_stamps')
    _lib.check(lib.cfb_debug_time_conv(*args), 'time_conv')
    _lib.check(lib.cfb_debug_set_stamps(None), 'set_stamps')
    torch.cuda.synchronize()
    s = stamps.cpu().tolist()
    cyc = s[9] - s[0]
    ns = s[15] - s[14]
    ghz = cyc / ns if ns > 0 else float('nan')
    print(f'--- N={n} {h}x{w} {cin}->{cout} k{k} xf={xf}: {base:.2f} us per launch back-to-back; CTA 0 lives {cyc} cycles = {ns} ns ({ghz:.2f} GHz)')
    for i in (1, 2, 10, 11, 12, 19, 20, 3, 4, 5, 6, 7, 16, 17, 18, 13, 8, 9):
        if s[i]:
            print(f'    {NAMES[i]:34s} +{(s[i] - s[0]) / ghz / 1e3:8.2f} us')


torch.zeros(1).cuda()
for shape in [(1, 16, 16, 512, 512, 1, False), (1, 16, 16, 64, 256, 1, False), (1, 16, 16, 512, 512, 3, True), (1, 32, 32, 256, 256, 3, True),
              (1, 128, 128, 128, 128, 3, True), (1, 256, 256, 128, 128, 3, True), (1, 512, 512, 64, 64, 3, True), (8, 256, 256, 128, 128, 3, True)]:
    run(*shape)
