"""Parameter tree of the hot path: names, shapes and block plans.

This is the *checkpoint contract* of the drop-in boundary: the names
and shapes below are exactly the ``state_dict`` keys produced by the reference classes
``VQAutoEncoder`` (/root/reference/basicsr/archs/vqgan_arch.py:326-382) and
``CodeFormer`` (/root/reference/basicsr/archs/codeformer_arch.py:160-212), so a
reference ``.pth`` loads with ``strict=True``.  Pure host-side metadata, no arithmetic.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import torch

# channels per feature size, codeformer_arch.py:194-201
FUSE_CHANNELS = {'16': 512, '32': 256, '64': 256, '128': 128, '256': 128, '512': 64}
FUSE_ENCODER_BLOCK = {'512': 2, '256': 5, '128': 8, '64': 11, '32': 14, '16': 18}      # codeformer_arch.py:204
FUSE_GENERATOR_BLOCK = {'16': 6, '32': 9, '64': 12, '128': 15, '256': 18, '512': 21}   # codeformer_arch.py:206


def encoder_plan(nf: int, ch_mult: Sequence[int], res_blocks: int, resolution: int,
                 attn_resolutions: Sequence[int], in_channels: int = 3, emb_dim: int = 256) -> List[tuple]:
    """(kind, cin, cout, out_res) per block -- mirrors Encoder.__init__ vqgan_arch.py:241-267."""
    plan = [('conv', in_channels, nf, resolution)]
    curr = resolution
    in_ch_mult = (1,) + tuple(ch_mult)
    cin = nf
    for i in range(len(ch_mult)):
        cin = nf * in_ch_mult[i]
        cout = nf * ch_mult[i]
        for _ in range(res_blocks):
            plan.append(('res', cin, cout, curr))
            cin = cout
            if curr in attn_resolutions:
                plan.append(('attn', cin, cin, curr))
        if i != len(ch_mult) - 1:
            curr //= 2
            plan.append(('down', cin, cin, curr))
    plan += [('res', cin, cin, curr), ('attn', cin, cin, curr), ('res', cin, cin, curr),
             ('norm', cin, cin, curr), ('conv', cin, emb_dim, curr)]
    return plan


def generator_plan(nf: int, ch_mult: Sequence[int], res_blocks: int, resolution: int,
                   attn_resolutions: Sequence[int], emb_dim: int = 256) -> List[tuple]:
    """(kind, cin, cout, out_res) per block -- mirrors Generator.__init__ vqgan_arch.py:290-316."""
    cin = nf * ch_mult[-1]
    curr = resolution // 2 ** (len(ch_mult) - 1)
    plan = [('conv', emb_dim, cin, curr), ('res', cin, cin, curr), ('attn', cin, cin, curr), ('res', cin, cin, curr)]
    for i in reversed(range(len(ch_mult))):
        cout = nf * ch_mult[i]
        for _ in range(res_blocks):
            plan.append(('res', cin, cout, curr))
            cin = cout
            if curr in attn_resolutions:
                plan.append(('attn', cin, cin, curr))
        if i != 0:
            curr *= 2
            plan.append(('up', cin, cin, curr))
    plan += [('norm', cin, cin, curr), ('conv', cin, 3, curr)]
    return plan


def _conv(spec, p, cin, cout, k):
    spec[p + '.weight'] = (cout, cin, k, k)
    spec[p + '.bias'] = (cout,)


def _norm(spec, p, c):
    spec[p + '.weight'] = (c,)
    spec[p + '.bias'] = (c,)


def _resblock(spec, p, cin, cout):
    _norm(spec, p + '.norm1', cin)
    _conv(spec, p + '.conv1', cin, cout, 3)
    _norm(spec, p + '.norm2', cout)
    _conv(spec, p + '.conv2', cout, cout, 3)
    if cin != cout:
        _conv(spec, p + '.conv_out', cin, cout, 1)


def _blocks(spec, prefix, plan):
    for i, (kind, cin, cout, _) in enumerate(plan):
        p = f'{prefix}.blocks.{i}'
        if kind == 'conv':
            _conv(spec, p, cin, cout, 3)
        elif kind == 'res':
            _resblock(spec, p, cin, cout)
        elif kind == 'attn':
            _norm(spec, p + '.norm', cin)
            for n in ('q', 'k', 'v', 'proj_out'):
                _conv(spec, f'{p}.{n}', cin, cin, 1)
        elif kind in ('down', 'up'):
            _conv(spec, p + '.conv', cin, cin, 3)
        elif kind == 'norm':
            _norm(spec, p, cin)


def vqae_spec(img_size=512, nf=64, ch_mult=(1, 2, 2, 4, 4, 8), res_blocks=2, attn_resolutions=(16,),
              codebook_size=1024, emb_dim=256) -> 'OrderedDict[str, tuple]':
    """state_dict keys/shapes of VQAutoEncoder(quantizer='nearest') in registration order."""
    spec: 'OrderedDict[str, tuple]' = OrderedDict()
    _blocks(spec, 'encoder', encoder_plan(nf, ch_mult, res_blocks, img_size, attn_resolutions, 3, emb_dim))
    spec['quantize.embedding.weight'] = (codebook_size, emb_dim)
    _blocks(spec, 'generator', generator_plan(nf, ch_mult, res_blocks, img_size, attn_resolutions, emb_dim))
    return spec


def codeformer_spec(dim_embd=512, n_head=8, n_layers=9, codebook_size=1024, latent_size=256,
                    connect_list=('32', '64', '128', '256')) -> 'OrderedDict[str, tuple]':
    """state_dict keys/shapes of CodeFormer (codeformer_arch.py:162-212) in registration order."""
    spec: 'OrderedDict[str, tuple]' = OrderedDict()
    spec['position_emb'] = (latent_size, dim_embd)       # direct parameter => first key of state_dict()
    spec.update(vqae_spec(512, 64, (1, 2, 2, 4, 4, 8), 2, (16,), codebook_size, 256))
    spec['feat_emb.weight'] = (dim_embd, 256)
    spec['feat_emb.bias'] = (dim_embd,)
    for l in range(n_layers):
        p = f'ft_layers.{l}'
        spec[p + '.self_attn.in_proj_weight'] = (3 * dim_embd, dim_embd)
        spec[p + '.self_attn.in_proj_bias'] = (3 * dim_embd,)
        spec[p + '.self_attn.out_proj.weight'] = (dim_embd, dim_embd)
        spec[p + '.self_attn.out_proj.bias'] = (dim_embd,)
        spec[p + '.linear1.weight'] = (2 * dim_embd, dim_embd)
        spec[p + '.linear1.bias'] = (2 * dim_embd,)
        spec[p + '.linear2.weight'] = (dim_embd, 2 * dim_embd)
        spec[p + '.linear2.bias'] = (dim_embd,)
        _norm(spec, p + '.norm1', dim_embd)
        _norm(spec, p + '.norm2', dim_embd)
    _norm(spec, 'idx_pred_layer.0', dim_embd)
    spec['idx_pred_layer.1.weight'] = (codebook_size, dim_embd)
    for s in connect_list:
        c = FUSE_CHANNELS[s]
        p = f'fuse_convs_dict.{s}'
        _resblock(spec, p + '.encode_enc', 2 * c, c)
        for br in ('scale', 'shift'):
            _conv(spec, f'{p}.{br}.0', c, c, 3)
            _conv(spec, f'{p}.{br}.2', c, c, 3)
    return spec


def random_state_dict(spec: 'OrderedDict[str, tuple]', seed: int = 1) -> 'OrderedDict[str, torch.Tensor]':
    """Seeded fp32 parameters that exercise every term (pretrained weights are not available
    offline; the reference's default init leaves position_emb = 0 and all norm affines =
    (1, 0)).  Deterministic for a given torch version: CPU generator, fixed
    key order.  matrices/filters U(+-1/sqrt(fan_in)); norm weights 1+0.1N; biases 0.1N;
    position_emb 0.02N; codebook N(0,1)."""
    g = torch.Generator(device='cpu')
    g.manual_seed(seed)
    sd: 'OrderedDict[str, torch.Tensor]' = OrderedDict()
    for name, shape in spec.items():
        if name == 'quantize.embedding.weight':
            t = torch.randn(shape, generator=g)
        elif name == 'position_emb':
            t = 0.02 * torch.randn(shape, generator=g)
        elif len(shape) >= 2:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            b = 1.0 / (fan_in ** 0.5)
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif name.endswith('.weight'):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            t = 0.1 * torch.randn(shape, generator=g)
        sd[name] = t.float().contiguous()
    return sd


def rrdbnet_spec(num_in_ch=3, num_out_ch=3, scale=4, num_feat=64, num_block=23, num_grow_ch=32) -> 'OrderedDict[str, tuple]':
    """state_dict keys/shapes of RRDBNet (basicsr/archs/rrdbnet_arch.py:86-101) in registration order."""
    spec: 'OrderedDict[str, tuple]' = OrderedDict()
    cin = num_in_ch * (4 if scale == 2 else (16 if scale == 1 else 1))
    _conv(spec, 'conv_first', cin, num_feat, 3)
    for b in range(num_block):
        for r in (1, 2, 3):
            for k in range(1, 6):
                _conv(spec, f'body.{b}.rdb{r}.conv{k}', num_feat + (k - 1) * num_grow_ch, num_feat if k == 5 else num_grow_ch, 3)
    for nm in ('conv_body', 'conv_up1', 'conv_up2', 'conv_hr'):
        _conv(spec, nm, num_feat, num_feat, 3)
    _conv(spec, 'conv_last', num_feat, num_out_ch, 3)
    return spec
