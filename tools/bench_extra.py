"""Secondary measurements (BASELINE.json configs[2] and configs[3]); one JSON line each.
  config 3: VectorQuantizer-only microbench, 32x256x16x16 vs 1024 codes (HBM roofline: 17.9 MB algorithmic)
  config 4: VQAutoEncoder.forward, batch 64 (580.59 GFLOP/face)
"""
import json
import sys

import torch

sys.path.insert(0, '.')
import codeformer_b200 as cb  # noqa: E402
from codeformer_b200 import spec as S  # noqa: E402

torch.set_grad_enabled(False)


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


g = torch.Generator().manual_seed(0)
E = torch.randn(1024, 256, generator=g)
z = torch.randn(32, 256, 16, 16, generator=g).cuda()
vq = cb.VectorQuantizer(1024, 256, 0.25)
vq.embedding.weight.data.copy_(E)
vq = vq.cuda()
ms = timed(lambda: vq(z, return_min_encodings=False), 50)
alg = 17.9e6
print(json.dumps({'config': 'VectorQuantizer 32x256x16x16 vs 1024 codes (configs[2])', 'ms': ms, 'vectors_per_s': 8192 / ms * 1e3,
                  'algorithmic_GBps': alg / ms / 1e6, 'hbm_peak_GBps': 6568.0, 'frac_of_hbm_peak': alg / ms / 1e6 / 6568.0,
                  'note': 'includes NCHW<->NHWC transposes, codebook split and the [T,K] dot-product round trip'}), flush=True)

net = cb.VQAutoEncoder(512, 64, [1, 2, 2, 4, 4, 8], 'nearest', 2, [16], 1024).cuda().eval()
net.load_state_dict(S.random_state_dict(S.vqae_spec(), 2))
x = torch.randn(64, 3, 512, 512, generator=g).clamp_(-1, 1).cuda()
ms = timed(lambda: net(x), 3, warm=2)
print(json.dumps({'config': 'VQAutoEncoder.forward batch 64 (configs[3])', 'ms_per_step': ms, 'faces_per_s': 64 / ms * 1e3,
                  'algorithmic_tflops': 64 * 580.59 / ms, 'launches': net.last_launch_count}), flush=True)
x1 = x[:1].contiguous()
cf = cb.CodeFormer().cuda().eval()
cf.load_state_dict(S.random_state_dict(S.codeformer_spec(), 1))
ms = timed(lambda: cf(x1, w=0.5, adain=True), 20)
print(json.dumps({'config': 'CodeFormer.forward batch 1 latency (configs[0] on GPU)', 'ms': ms, 'launches': cf.last_launch_count}), flush=True)
