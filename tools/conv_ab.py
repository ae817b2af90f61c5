#!/usr/bin/env python
"""CUDA-event time of single tcgen05 conv launches (cfb_debug_time_conv) over the shapes of one forward -- the cheap A/B
of two library builds on one box: CFB_LIB=<path to another libcfb200.so> python tools/conv_ab.py [--small]."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402

SHAPES = [  # N, H, Cin, Cout, fused transform
    (8, 256, 128, 128, True), (8, 256, 128, 128, False), (8, 512, 64, 64, True), (8, 128, 256, 256, True),
    (8, 64, 256, 256, True), (8, 32, 512, 512, True), (8, 16, 512, 512, True),
    (1, 512, 64, 64, True), (1, 256, 128, 128, True), (1, 128, 256, 256, True), (1, 64, 256, 256, True),
    (1, 32, 512, 512, True), (1, 16, 512, 512, True)]


def main():
    torch.zeros(1).cuda()
    tag = os.path.basename(os.environ.get('CFB_LIB', 'default'))
    res = []
    for (n, h, ci, co, xf) in SHAPES:
        if '--small' in sys.argv and n != 1:
            continue
        ms = min(bench.time_conv_kernel(torch, n, h, ci, co, xf, reps=20) for _ in range(2))
        res.append(f'{n}x{h}^2 {ci}->{co}{" xf" if xf else " raw"}: {ms * 1e3:.1f}')
    print(f'[{tag} pdl={os.environ.get("CFB_PDL", "1")}] us/launch  ' + ' | '.join(res), flush=True)


if __name__ == '__main__':
    main()
