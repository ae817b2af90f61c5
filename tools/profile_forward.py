"""One CodeFormer.forward under the CUDA profiler range, for ncu:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
       python tools/profile_forward.py --batch 8
"""
import argparse
import sys

import torch

sys.path.insert(0, '.')
import codeformer_b200 as cb  # noqa: E402
from codeformer_b200 import spec as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=8)
ap.add_argument('--engine', default='auto')
ap.add_argument('--vqae', action='store_true')
args = ap.parse_args()
torch.set_grad_enabled(False)
if args.vqae:
    net = cb.VQAutoEncoder(512, 64, [1, 2, 2, 4, 4, 8], 'nearest', 2, [16], 1024).cuda().eval()
    net.load_state_dict(S.random_state_dict(S.vqae_spec(), 2))
else:
    net = cb.CodeFormer().cuda().eval()
    net.load_state_dict(S.random_state_dict(S.codeformer_spec(), 1))
net.set_engine(args.engine)
x = torch.randn(args.batch, 3, 512, 512, generator=torch.Generator().manual_seed(0)).clamp_(-1, 1).cuda()
run = (lambda: net(x)) if args.vqae else (lambda: net(x, w=0.5, adain=True))
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('launches', net.last_launch_count)
