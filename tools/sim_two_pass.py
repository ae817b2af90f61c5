"""CPU experiment: what does dropping the A_lo x B_hi pass cost in the generator / fuse convs?

Runs the oracle forward twice on golden face 0 (seeded weights of the bench): once plain fp32, once with the INPUT of every
generator / fuse conv rounded to fp16 (= the A_hi operand alone; weights keep hi+lo).  Prints max-abs / rms differences of
`out`.  Test infrastructure only (imports oracle/)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from oracle import codeformer_oracle as O
from codeformer_b200 import spec as S

torch.set_grad_enabled(False)
torch.set_num_threads(os.cpu_count())
sd = S.random_state_dict(S.codeformer_spec(), 1)
faces = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'faces.npz'))['faces']
x = torch.from_numpy(faces[:2].astype(np.float32) / 255.0).permute(0, 3, 1, 2)[:, [2, 1, 0]]
x = (x - 0.5) / 0.5
ref, logits, _ = O.codeformer_forward(sd, x, w=0.5, adain_on=True)
orig = O.conv
mode = sys.argv[1] if len(sys.argv) > 1 else 'act'
which = sys.argv[2] if len(sys.argv) > 2 else 'gen'
def conv2(sd_, p, x_, stride=1, padding=1):
    hit = p.startswith('generator') or p.startswith('fuse_convs_dict') if which == 'gen' else True
    if hit:
        if mode == 'act':
            x_ = x_.half().float()
        elif mode == 'wgt':
            sd_ = dict(sd_); sd_[p + '.weight'] = sd_[p + '.weight'].half().float()
    return orig(sd_, p, x_, stride, padding)
O.conv = conv2
out, logits2, _ = O.codeformer_forward(sd, x, w=0.5, adain_on=True)
d = (out - ref).abs()
print(f'mode={mode} which={which}: out range [{ref.min():.3f},{ref.max():.3f}] rms {ref.pow(2).mean().sqrt():.3f}; '
      f'max abs diff {d.max():.3e}, rms diff {d.pow(2).mean().sqrt():.3e}; idx equal {bool((logits.argmax(2) == logits2.argmax(2)).all())}')
