"""CodeFormer.forward throughput vs batch size (is the step limited by HBM traffic of the 256^2/512^2 layers,
i.e. would L2-resident sub-batches pay?).  One JSON line per batch size."""
import json
import sys

import torch

sys.path.insert(0, '.')
import codeformer_b200 as cb  # noqa: E402
from codeformer_b200 import spec as S  # noqa: E402

torch.set_grad_enabled(False)
cf = cb.CodeFormer().cuda().eval()
cf.load_state_dict(S.random_state_dict(S.codeformer_spec(), 1))
g = torch.Generator().manual_seed(0)
x = torch.randn(32, 3, 512, 512, generator=g).clamp_(-1, 1).cuda()
for B in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 6, 8, 16, 32]:
    xb = x[:B].contiguous()
    iters = max(3, 64 // B)
    for _ in range(3):
        cf(xb, w=0.5, adain=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        cf(xb, w=0.5, adain=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(json.dumps({'batch': B, 'ms': round(ms, 3), 'faces_per_s': round(B / ms * 1e3, 1)}), flush=True)
