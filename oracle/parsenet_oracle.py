"""CPU oracle for ParseNet -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (only tests/, smoke() and bench.py's CPU legs import it).

Functional fp32 restatement of /root/reference/facelib/parsing/parsenet.py (ConvLayer :72-105, ResidualBlock :108-137,
ParseNet.forward :188-194) with the torch operators the reference calls; eval-mode BatchNorm.  Pinned by
tests/golden/parsenet.npz, written by oracle/gen_golden.py from the UNMODIFIED reference class.
"""
import torch
import torch.nn.functional as F


def conv_layer(sd, p, x, scale='none', norm=False, lrelu=False):
    """ConvLayer.forward  parsenet.py:98-105: [nearest x2] -> ReflectionPad2d(1) -> conv3x3 (stride 2 if 'down') -> [BN] -> [LeakyReLU 0.2]"""
    if scale == 'up':
        x = F.interpolate(x, scale_factor=2, mode='nearest')
    x = F.pad(x, (1, 1, 1, 1), mode='reflect')
    x = F.conv2d(x, sd[p + '.conv2d.weight'], sd.get(p + '.conv2d.bias'), stride=2 if scale == 'down' else 1)
    if norm:
        q = p + '.norm.norm.'
        x = F.batch_norm(x, sd[q + 'running_mean'], sd[q + 'running_var'], sd[q + 'weight'], sd[q + 'bias'], training=False, eps=1e-5)
    if lrelu:
        x = F.leaky_relu(x, 0.2)
    return x


def residual_block(sd, p, x, kind):
    """ResidualBlock.forward  parsenet.py:131-137 with scale_config_dict :124-125"""
    conf = {'down': ('none', 'down'), 'up': ('up', 'none'), 'none': ('none', 'none')}[kind]
    identity = x if (p + '.shortcut_func.conv2d.weight') not in sd else conv_layer(sd, p + '.shortcut_func', x, kind)
    res = conv_layer(sd, p + '.conv1', x, conf[0], norm=True, lrelu=True)
    res = conv_layer(sd, p + '.conv2', res, conf[1], norm=True)
    return identity + res


def parsenet_forward(sd, x, plan):
    """plan: codeformer_b200.parsing.parsenet_plan(...)[0] -- (prefix, kind, cin, cout) per ResidualBlock."""
    feat = conv_layer(sd, 'encoder.0', x)
    for p, kind, _, _ in plan:
        if p.startswith('encoder'):
            feat = residual_block(sd, p, feat, kind)
    y = feat
    for p, kind, _, _ in plan:
        if p.startswith('body'):
            y = residual_block(sd, p, y, kind)
    y = feat + y
    for p, kind, _, _ in plan:
        if p.startswith('decoder'):
            y = residual_block(sd, p, y, kind)
    return conv_layer(sd, 'out_mask_conv', y), conv_layer(sd, 'out_img_conv', y)
