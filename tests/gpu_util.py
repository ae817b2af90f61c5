"""ctypes-level helpers for the -m gpu parity tests (they call through the C ABI of include/cfb200.h)."""
import ctypes

import torch

from codeformer_b200 import _lib


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def nhwc(x):
    """NCHW cpu/cuda tensor -> contiguous NHWC cuda tensor"""
    return x.permute(0, 2, 3, 1).contiguous().cuda()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def conv2d(x_nchw, weight, bias, mode=0, in_scale=None, in_shift=None, in_act=0, residual=None, out_act=0, engine=1):
    """cfb_conv2d_nhwc on an NCHW host tensor; returns NCHW cuda tensor."""
    lib = _lib.load()
    N, Cin, H, W = x_nchw.shape
    Cout, _, k, _ = weight.shape
    Ho = H // 2 if mode == 1 else (H * 2 if mode == 2 else H)
    Wo = W // 2 if mode == 1 else (W * 2 if mode == 2 else W)
    xin = nhwc(x_nchw)
    w = weight.contiguous().cuda()
    b = None if bias is None else bias.contiguous().cuda()
    out = torch.empty((N, Ho, Wo, Cout), device='cuda')
    sc = None if in_scale is None else in_scale.contiguous().cuda()
    sh = None if in_shift is None else in_shift.contiguous().cuda()
    res = None if residual is None else nhwc(residual)
    wsb = lib.cfb_conv2d_workspace_bytes(N, H, W, Cin, Cout, k, mode)
    ws = torch.empty(int(wsb), dtype=torch.uint8, device='cuda')
    _lib.check(lib.cfb_conv2d_nhwc(_lib.ptr(xin), _lib.ptr(w), _lib.ptr(b), _lib.ptr(out), N, H, W, Cin, Cout, k, mode,
                                   _lib.ptr(sc), _lib.ptr(sh), in_act, _lib.ptr(res), out_act, engine,
                                   _lib.ptr(ws), wsb, stream()), 'cfb_conv2d_nhwc')
    torch.cuda.synchronize()
    return nchw(out)


def gn_coef(x_nchw, gamma, beta, groups=32, eps=1e-6):
    lib = _lib.load()
    N, C, H, W = x_nchw.shape
    xin = nhwc(x_nchw)
    scale = torch.empty((N, C), device='cuda')
    shift = torch.empty((N, C), device='cuda')
    wsb = lib.cfb_gn_workspace_bytes(N, H * W, C)
    ws = torch.empty(int(wsb), dtype=torch.uint8, device='cuda')
    g_d, b_d = gamma.cuda(), beta.cuda()        # keep alive: ptr() of a temporary dangles once it is freed
    _lib.check(lib.cfb_group_norm_coef(_lib.ptr(xin), _lib.ptr(g_d), _lib.ptr(b_d), _lib.ptr(scale),
                                       _lib.ptr(shift), N, H * W, C, groups, eps, _lib.ptr(ws), wsb, stream()),
               'cfb_group_norm_coef')
    torch.cuda.synchronize()
    return xin, scale, shift
