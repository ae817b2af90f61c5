"""-m gpu: the caller-side networks of SURVEY.md section 8 rows f3 / f4 on the tcgen05 engine (generalised fused-transform conv:
any H x W, reflect / replicate padding, dense-buffer placement, second residual, stride 2 by subsampling) against torch CPU
fp32, the CPU oracle and the golden vectors of the UNMODIFIED reference.  Tolerances are written at each check."""
import ctypes
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import codeformer_b200 as cb
from codeformer_b200 import _lib
from codeformer_b200 import spec as S
from tests.util import golden, maxabs

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def _gen_conv(x_nchw, w, b, up=0, pad_mode=0, sub=0, act=0, res=None, res2=None, post=1.0, in_pitch=None, out_pitch=None, out_c0=0):
    """cfb_conv2d_gen_nhwc on host tensors; returns the NCHW result (the [out_c0, out_c0+cout) slice of the destination)."""
    lib = _lib.load()
    N, Cin, H, W = x_nchw.shape
    Cout = w.shape[0]
    in_pitch = in_pitch or Cin
    xin = torch.full((N, H, W, in_pitch), 7.25)                     # channels beyond cin hold finite junk (they meet zero weights)
    xin[..., :Cin] = x_nchw.permute(0, 2, 3, 1)
    xin = xin.cuda()
    Ho, Wo = (2 * H, 2 * W) if up else ((H // 2, W // 2) if sub else (H, W))
    out_pitch = out_pitch or Cout
    out = torch.full((N, Ho, Wo, out_pitch), -3.0, device='cuda')
    wd, bd = w.contiguous().cuda(), (None if b is None else b.contiguous().cuda())
    rd = None if res is None else res.permute(0, 2, 3, 1).contiguous().cuda()
    r2d = None if res2 is None else res2.permute(0, 2, 3, 1).contiguous().cuda()
    wsb = lib.cfb_conv2d_gen_workspace_bytes(Cin, Cout)
    ws = torch.empty(int(wsb), dtype=torch.uint8, device='cuda')
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.cfb_conv2d_gen_nhwc(_lib.ptr(xin), in_pitch, _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(out), out_pitch, out_c0, N, H, W,
                                       Cin, Cout, up, pad_mode, sub, act, _lib.ptr(rd), Cout, _lib.ptr(r2d), Cout, post,
                                       _lib.ptr(ws), wsb, st), 'cfb_conv2d_gen_nhwc')
    torch.cuda.synchronize()
    cb.check_async_status()
    full = out.cpu()
    if out_pitch != Cout:                                            # the rest of the destination buffer is untouched
        mask = torch.ones(out_pitch, dtype=torch.bool)
        mask[out_c0:out_c0 + Cout] = False
        assert bool((full[..., mask] == -3.0).all())
    return full[..., out_c0:out_c0 + Cout].permute(0, 3, 1, 2).contiguous()


PADS = {0: 'constant', 1: 'reflect', 2: 'replicate'}
# (N, Cin, Cout, H, W, pad_mode)
GEN_CASES = [(1, 64, 32, 20, 28, 0), (2, 96, 32, 17, 23, 0), (1, 160, 32, 40, 9, 0), (1, 192, 64, 33, 31, 0),
             (1, 64, 64, 24, 40, 1), (2, 128, 128, 19, 21, 1), (1, 256, 128, 16, 16, 1), (1, 64, 64, 50, 7, 2)]


@pytest.mark.parametrize('case', GEN_CASES)
def test_gen_conv_ragged_sizes_and_padding(case):
    """3x3 stride-1 conv, any H x W, zero / reflect / replicate padding, real channel counts that are not multiples of 64
    (dense-block windows), LeakyReLU, written into a slice of a wider buffer."""
    N, Cin, Cout, H, W, pm = case
    x = _rand(N, Cin, H, W, seed=1)
    w = _rand(Cout, Cin, 3, 3, seed=2, scale=1 / math.sqrt(9 * Cin))
    b = _rand(Cout, seed=3, scale=0.1)
    ref = F.leaky_relu(F.conv2d(F.pad(x, (1, 1, 1, 1), mode=PADS[pm]), w, b), 0.2)
    in_pitch = (Cin + 63) // 64 * 64 + (64 if Cin % 64 else 0)
    out = _gen_conv(x, w, b, pad_mode=pm, act=1, in_pitch=in_pitch, out_pitch=Cout + 32, out_c0=16)
    assert maxabs(out, ref) < 6e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize('pm', [0, 1])
def test_gen_conv_stride2_and_upsample_and_residuals(pm):
    N, C, H, W = 1, 64, 26, 18
    x = _rand(N, C, H, W, seed=4)
    w = _rand(128, C, 3, 3, seed=5, scale=1 / math.sqrt(9 * C))
    b = _rand(128, seed=6, scale=0.1)
    # stride 2, pad 1 (ParseNet 'down', parsenet.py:88,101-102): the even positions of the stride-1 result
    ref = F.conv2d(F.pad(x, (1, 1, 1, 1), mode=PADS[pm]), w, b, stride=2)
    out = _gen_conv(x, w, b, pad_mode=pm, sub=1)
    assert out.shape == ref.shape and maxabs(out, ref) < 6e-5 * float(ref.abs().max())
    # nearest x2 then (reflect-)padded conv (ParseNet 'up', parsenet.py:92-93; RRDBNet conv_up, rrdbnet_arch.py:116-117)
    up = F.interpolate(x, scale_factor=2, mode='nearest')
    ref = F.leaky_relu(F.conv2d(F.pad(up, (1, 1, 1, 1), mode=PADS[pm]), w, b), 0.2)
    out = _gen_conv(x, w, b, up=1, pad_mode=2 if pm == 1 else 0, act=1)        # reflect after nearest-x2 == replicate before it
    assert out.shape == ref.shape and maxabs(out, ref) < 6e-5 * float(ref.abs().max())
    # two residuals: (conv + bias + r1) * 0.2 + r2   (rrdbnet_arch.py:40,63)
    w2 = _rand(64, C, 3, 3, seed=7, scale=1 / math.sqrt(9 * C))
    b2 = _rand(64, seed=8, scale=0.1)
    r1, r2 = _rand(N, 64, H, W, seed=9), _rand(N, 64, H, W, seed=10)
    ref = (F.conv2d(x, w2, b2, padding=1) + r1) * 0.2 + r2
    out = _gen_conv(x, w2, b2, res=r1, res2=r2, post=0.2)
    assert maxabs(out, ref) < 6e-5 * float(ref.abs().max())


@pytest.mark.parametrize('case', ['s2', 's4'])
def test_rrdbnet_vs_reference_golden(case):
    """RRDBNet.forward (23 RRDBs) against the output of the UNMODIFIED reference class on CPU fp32: <= 1e-3 max-abs
    (north_star tolerance for fp32 outputs; observed ~1e-5 on outputs of magnitude ~1)."""
    from oracle import gen_golden as GG
    scale, sd, x = GG.rrdb_inputs(case)
    net = cb.ARCH_REGISTRY.get('RRDBNet')(3, 3, scale=scale, num_feat=64, num_block=23, num_grow_ch=32).cuda().eval()
    net.load_state_dict(sd, strict=True)
    out = net(x.cuda())
    torch.cuda.synchronize()
    cb.check_async_status()
    ref = golden('rrdbnet.npz')[case + '_out']
    err = maxabs(out.cpu(), ref)
    print(f'rrdbnet {case}: max-abs {err:.3e} (|out|max {float(np.abs(ref).max()):.2f})')
    assert out.shape == ref.shape and err < 1e-3
    assert torch.equal(net(x.cuda()), out), 'deterministic'


def test_rrdbnet_vs_oracle_odd_sizes_and_batch_invariance():
    from oracle import rrdbnet_oracle as RO
    sd = S.random_state_dict(S.rrdbnet_spec(3, 3, 4, 64, 3, 32), 9)
    net = cb.RRDBNet(3, 3, scale=4, num_block=3).cuda().eval()
    net.load_state_dict(sd)
    x = torch.rand(3, 3, 37, 53, generator=torch.Generator().manual_seed(1))
    out = net(x.cuda())
    ref = RO.rrdbnet_forward(sd, x, scale=4, num_block=3)
    assert maxabs(out.cpu(), ref) < 1e-4
    assert torch.equal(net(x[1:2].cuda())[0], out[1]), 'images must not interact'
    e = net(torch.empty(0, 3, 8, 8, device='cuda'))
    assert e.shape == (0, 3, 32, 32)
    with pytest.raises(RuntimeError):
        net(x.cuda().half())


def test_realesrganer_tiles_on_the_gpu_equal_the_cpu_oracle_pipeline():
    """RealESRGANer.enhance with the tile loop on the GPU model vs the same front-end around the CPU oracle: uint8 images equal
    up to one LSB at a handful of rounding boundaries."""
    from oracle import rrdbnet_oracle as RO
    sd = S.random_state_dict(S.rrdbnet_spec(3, 3, 2, 64, 2, 32), 11)
    net = cb.RRDBNet(3, 3, scale=2, num_block=2)
    net.load_state_dict(sd)

    class Oracle(torch.nn.Module):
        def forward(self, x):
            return RO.rrdbnet_forward(sd, x, scale=2, num_block=2)
    img = np.random.default_rng(3).integers(0, 256, (75, 61, 3), dtype=np.uint8)
    a, _ = cb.RealESRGANer(scale=2, model=net, tile=32, tile_pad=8, pre_pad=0, device='cuda').enhance(img, outscale=2)
    b, _ = cb.RealESRGANer(scale=2, model=Oracle(), tile=32, tile_pad=8, pre_pad=0, device='cpu').enhance(img, outscale=2)
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    assert a.shape == (150, 122, 3) and d.max() <= 1 and (d > 0).mean() < 1e-3


def test_parsenet_vs_reference_golden():
    """ParseNet(512, 512, parsing_ch=19) on committed face 0 against the UNMODIFIED reference on CPU fp32: logits <= 2e-4
    max-abs (|logit|max 0.38), argmax classes identical wherever the reference's top-1/top-2 margin exceeds 1e-3."""
    from oracle import gen_golden as GG
    sd, x = GG.parsenet_inputs()
    net = cb.ParseNet(in_size=512, out_size=512, parsing_ch=19)
    net.load_state_dict(sd, strict=True)
    net = net.eval().cuda()
    mask, img = net(x.cuda())
    torch.cuda.synchronize()
    cb.check_async_status()
    g = golden('parsenet.npz')
    e_m, e_i = maxabs(mask[..., ::4, ::4].cpu(), g['mask_s4']), maxabs(img[..., ::8, ::8].cpu(), g['img_s8'])
    print(f'parsenet: mask max-abs {e_m:.3e} img {e_i:.3e}')
    assert mask.shape == (1, 19, 512, 512) and img.shape == (1, 3, 512, 512)
    assert e_m < 2e-4 and e_i < 2e-4
    cls, fm = cb.face_parse_mask(mask)
    sure = torch.from_numpy(g['margin'].astype(np.float32)) > 1e-3
    assert torch.equal(cls.cpu()[sure], torch.from_numpy(g['classes'])[sure]), 'class indices'
    assert torch.equal(cls.cpu().long(), mask.argmax(1).cpu()), 'face_parse_mask == torch.argmax on the same logits'
    lut = torch.tensor([0] + [255] * 13 + [0, 255, 0, 0, 0], dtype=torch.uint8)              # face_restoration_helper.py:465
    assert torch.equal(fm.cpu(), lut[cls.cpu().long()])
    assert torch.equal(net(x.cuda())[0], mask), 'deterministic'


@pytest.mark.parametrize('size,batch', [(64, 3), (128, 1)])
def test_parsenet_small_configs_vs_oracle(size, batch):
    """Other constructor arguments (1 and 2 down/up steps) and a batch, against the CPU oracle; faces must not interact."""
    from codeformer_b200 import parsing as P
    from oracle import parsenet_oracle as PO
    sd = P.random_parsenet_state_dict(P.parsenet_spec(size, size), 5)
    net = cb.ParseNet(in_size=size, out_size=size)
    net.load_state_dict(sd, strict=True)
    net = net.eval().cuda()
    x = torch.rand(batch, 3, size, size, generator=torch.Generator().manual_seed(2)) * 2 - 1
    mask, img = net(x.cuda())
    rm, ri = PO.parsenet_forward(sd, x, P.parsenet_plan(size, size)[0])
    assert maxabs(mask.cpu(), rm) < 2e-4 * max(1.0, float(rm.abs().max())) and maxabs(img.cpu(), ri) < 2e-4 * max(1.0, float(ri.abs().max()))
    assert torch.equal(net(x[:1].cuda())[0][0], mask[0])
