#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "vq" --timeout 300 > gpurun_out/t_vq.log 2>&1; echo "vq tests rc=$?"; tail -15 gpurun_out/t_vq.log | cut -c1-300
timeout 300 python - <<'PY' 2>&1 | tail -5
import torch, os, sys
sys.path.insert(0, '.')
import codeformer_b200 as cb
g = torch.Generator().manual_seed(0)
E = torch.randn(1024, 256, generator=g); z = torch.randn(32, 256, 16, 16, generator=g).cuda()
vq = cb.VectorQuantizer(1024, 256, 0.25); vq.embedding.weight.data.copy_(E); vq = vq.cuda()
def t(fn, it=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
for fused in ('1', '0'):
    os.environ['CFB_VQ_FUSED'] = fused
    for graphs in (True, False):
        vq.__dict__.pop('_cfb_vq', None); vq.vq_graphs = graphs
        print(f'fused={fused} graphs={graphs}: {t(lambda: vq(z, return_min_encodings=False)):.1f} us per call (back-to-back, incl. copies)')
# kernel alone through the C ABI
from codeformer_b200 import _lib
import ctypes
lib = _lib.load()
os.environ['CFB_VQ_FUSED'] = '1'
Ed = vq.embedding.weight.detach().contiguous()
prep = torch.empty(int(lib.cfb_vq_prepared_bytes(1024, 256)), dtype=torch.uint8, device='cuda')
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
_lib.check(lib.cfb_vq_prepare(_lib.ptr(Ed), 1024, 256, _lib.ptr(prep), prep.numel(), st))
zq = torch.empty_like(z); idx = torch.empty((8192, 1), dtype=torch.int64, device='cuda'); stats = torch.empty(4, device='cuda')
ws = torch.empty(int(lib.cfb_vq_fast_workspace_bytes(32, 256, 256, 1024)), dtype=torch.uint8, device='cuda')
call = lambda: _lib.check(lib.cfb_vq_nearest_fast(_lib.ptr(z), _lib.ptr(Ed), _lib.ptr(prep), 32, 16, 16, 256, 1024, 0.25, _lib.ptr(zq), _lib.ptr(idx), _lib.ptr(stats), None, _lib.ptr(ws), ws.numel(), st))
print(f'C ABI, one kernel, back-to-back launches: {t(call):.1f} us per call')
os.environ['CFB_VQ_TIMING'] = '1'
call(); torch.cuda.synchronize()
base = (ws.data_ptr() + 1023) // 1024 * 1024 - ws.data_ptr() + 4096
stamps = ws[base:base + 64 * 6 * 8].view(torch.int64).view(64, 6).cpu().double()
d = stamps - stamps[:, :1]
print('phase stamps (cycles from CTA start; mean over 64 CTAs): after phase0+cluster sync %.0f, producers done %.0f, epilogue done %.0f, phase 2 done %.0f, end %.0f'
      % tuple(d[:, k].mean().item() for k in (1, 2, 3, 4, 5)))
print('  max over CTAs: %s' % [int(d[:, k].max().item()) for k in (1, 2, 3, 4, 5)])
os.environ['CFB_VQ_TIMING'] = '0'

gr = torch.cuda.CUDAGraph()
call(); torch.cuda.synchronize()
with torch.cuda.graph(gr):
    for _ in range(20): call()
print(f'C ABI, one kernel, 20 calls per graph replay: {t(gr.replay, 50) / 20:.1f} us per call')
PY
