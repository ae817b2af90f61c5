"""Shared helpers for the test-suite (tests may use the oracle; the product never does)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def faces_input(sel=slice(None)):
    """Committed 512x512 RGB u8 faces -> f32 NCHW in [-1,1] (arithmetic of inference_codeformer.py:199-200)."""
    f = golden('faces.npz')['faces'][sel]
    t = torch.from_numpy(f.astype(np.float32) / 255.).permute(0, 3, 1, 2).contiguous()
    return (t - 0.5) / 0.5


def vq_micro_inputs(case):
    """Config-3 inputs: identical code to oracle/gen_golden.py."""
    g = torch.Generator().manual_seed(0)
    E = torch.randn(1024, 256, generator=g)
    if case == 'B':
        z = torch.randn(32, 256, 16, 16, generator=g)
    else:
        idx = torch.randint(0, 1024, (32 * 256,), generator=g)
        z = (E[idx] + 0.3 * torch.randn(32 * 256, 256, generator=g)).view(32, 16, 16, 256).permute(0, 3, 1, 2).contiguous()
    return E, z


def maxabs(a, b):
    return float((torch.as_tensor(a).float() - torch.as_tensor(b).float()).abs().max())
