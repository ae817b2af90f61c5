"""CPU oracle for RRDBNet -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (only tests/, smoke() and bench.py's CPU legs import it).

Functional (state dict in, tensor out) fp32 restatement of /root/reference/basicsr/archs/rrdbnet_arch.py with the torch
operators the reference calls at the cited lines (its arithmetic lives in PyTorch, not under /root/reference).  Pinned by
tests/golden/rrdbnet.npz, which oracle/gen_golden.py writes from the UNMODIFIED reference class (tests/test_oracle_aux.py).
"""
import torch
import torch.nn.functional as F


def pixel_unshuffle(x, scale):
    """basicsr/archs/arch_util.py:190-206"""
    b, c, hh, hw = x.shape
    assert hh % scale == 0 and hw % scale == 0
    h, w = hh // scale, hw // scale
    return x.view(b, c, h, scale, w, scale).permute(0, 1, 3, 5, 2, 4).reshape(b, c * scale * scale, h, w)


def _conv(sd, name, x):
    return F.conv2d(x, sd[name + '.weight'], sd[name + '.bias'], stride=1, padding=1)


def rdb(sd, p, x):
    """ResidualDenseBlock.forward  rrdbnet_arch.py:33-40"""
    lrelu = lambda t: F.leaky_relu(t, 0.2)
    x1 = lrelu(_conv(sd, p + '.conv1', x))
    x2 = lrelu(_conv(sd, p + '.conv2', torch.cat((x, x1), 1)))
    x3 = lrelu(_conv(sd, p + '.conv3', torch.cat((x, x1, x2), 1)))
    x4 = lrelu(_conv(sd, p + '.conv4', torch.cat((x, x1, x2, x3), 1)))
    x5 = _conv(sd, p + '.conv5', torch.cat((x, x1, x2, x3, x4), 1))
    return x5 * 0.2 + x


def rrdbnet_forward(sd, x, scale=4, num_block=23):
    """RRDBNet.forward  rrdbnet_arch.py:103-119"""
    feat = pixel_unshuffle(x, 2) if scale == 2 else (pixel_unshuffle(x, 4) if scale == 1 else x)
    feat = _conv(sd, 'conv_first', feat)
    body = feat
    for b in range(num_block):                                   # RRDB.forward  :58-63
        out = rdb(sd, f'body.{b}.rdb1', body)
        out = rdb(sd, f'body.{b}.rdb2', out)
        out = rdb(sd, f'body.{b}.rdb3', out)
        body = out * 0.2 + body
    feat = feat + _conv(sd, 'conv_body', body)
    feat = F.leaky_relu(_conv(sd, 'conv_up1', F.interpolate(feat, scale_factor=2, mode='nearest')), 0.2)
    feat = F.leaky_relu(_conv(sd, 'conv_up2', F.interpolate(feat, scale_factor=2, mode='nearest')), 0.2)
    return _conv(sd, 'conv_last', F.leaky_relu(_conv(sd, 'conv_hr', feat), 0.2))
