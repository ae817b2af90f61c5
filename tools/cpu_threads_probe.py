"""How does the oracle (reference CPU path) scale with torch threads on this box?  B=1, one warm + one timed pass each."""
import os
import sys
import time

import torch

sys.path.insert(0, '.')
from codeformer_b200 import spec as S  # noqa: E402
from oracle import codeformer_oracle as O  # noqa: E402

torch.set_grad_enabled(False)
sd = S.random_state_dict(S.codeformer_spec(), 1)
x = torch.randn(1, 3, 512, 512).clamp_(-1, 1)
print('cpu_count', os.cpu_count(), flush=True)
for nt in (8, 16, 32, 64, os.cpu_count()):
    torch.set_num_threads(nt)
    O.codeformer_forward(sd, x, w=0.5, adain_on=True)
    t = time.perf_counter()
    O.codeformer_forward(sd, x, w=0.5, adain_on=True)
    dt = time.perf_counter() - t
    print(f'threads={nt}: {dt:.2f} s/face -> {1 / dt:.3f} faces/s', flush=True)
