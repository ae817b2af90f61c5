"""RRDBNet + RealESRGANer on libcfb200.

Mirrors, for the caller,
    RRDBNet(num_in_ch, num_out_ch, scale, num_feat, num_block, num_grow_ch)   /root/reference/basicsr/archs/rrdbnet_arch.py:67-120
    RealESRGANer(scale, model_path, model, tile, tile_pad, pre_pad, half).enhance(img, outscale)
                                                                               /root/reference/basicsr/utils/realesrgan_utils.py:14-250
as they are built by ``set_realesrgan()`` (/root/reference/inference_codeformer.py:36-61) and called on the background image
and on restored faces.  The network's arithmetic is ``cfb_rrdb_forward`` (CUDA, include/cfb200.h); this module owns the
parameters (same state-dict keys, strict load of a reference checkpoint) and the image plumbing around the model call --
colour handling, reflect pre/mod padding, the tile loop -- written against torch tensors on the device.  No CPU fallback.
"""
import ctypes
import math
import threading

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import spec as S
from .registry import ARCH_REGISTRY


@ARCH_REGISTRY.register()
class RRDBNet(nn.Module):
    """Parameters of the reference's RRDBNet (identical ``state_dict``) with ``forward`` on the tcgen05 conv engine."""

    def __init__(self, num_in_ch, num_out_ch, scale=4, num_feat=64, num_block=23, num_grow_ch=32):
        super().__init__()
        if num_feat != 64 or num_grow_ch != 32:
            raise NotImplementedError('codeformer_b200 builds RRDBNet for num_feat=64, num_grow_ch=32 (the RealESRGAN models)')
        self.scale, self.num_in_ch, self.num_out_ch = scale, num_in_ch, num_out_ch
        self.num_feat, self.num_block, self.num_grow_ch = num_feat, num_block, num_grow_ch
        g = torch.Generator().manual_seed(0)
        params = {}
        for name, shape in S.rrdbnet_spec(num_in_ch, num_out_ch, scale, num_feat, num_block, num_grow_ch).items():
            if name.endswith('.weight'):                       # kaiming-normal * 0.1 like default_init_weights (arch_util.py:18-36)
                fan_in = shape[1] * shape[2] * shape[3]
                t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in) * 0.1
            else:
                t = torch.zeros(shape)
            params[name] = t
        # nested modules so that state_dict() yields the reference's dotted names
        self._register_tree(params)
        object.__setattr__(self, '_lock', threading.Lock())
        object.__setattr__(self, '_net', None)
        object.__setattr__(self, '_sig', None)
        object.__setattr__(self, '_keep', None)
        object.__setattr__(self, '_ws', None)

    def _register_tree(self, params):
        for name, t in params.items():
            mod = self
            parts = name.split('.')
            for p in parts[:-1]:
                if not hasattr(mod, p):
                    mod.add_module(p, nn.Module())
                mod = getattr(mod, p)
            mod.register_parameter(parts[-1], nn.Parameter(t))

    def _prepare(self, device):
        lib = _lib.load()
        params = list(self.state_dict(keep_vars=True).items())
        sig = tuple((k, v.data_ptr(), v._version, str(v.device)) for k, v in params)
        if self._net is not None and sig == self._sig:
            return
        if self._net is None:
            h = lib.cfb_rrdb_create(self.num_in_ch, self.num_out_ch, self.scale, self.num_feat, self.num_block, self.num_grow_ch)
            if not h:
                _lib.check(1, 'cfb_rrdb_create')
            object.__setattr__(self, '_net', ctypes.c_void_p(h))
        keep = []
        for k, v in params:
            if v.device != device:
                raise RuntimeError(f'parameter {k} is on {v.device} but the input is on {device}; call net.to(device)')
            t = v.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            keep.append(t)
            _lib.check(lib.cfb_rrdb_set_param(self._net, k.encode(), _lib.ptr(t), t.numel()), 'cfb_rrdb_set_param')
        _lib.check(lib.cfb_rrdb_prepare(self._net, ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)), 'cfb_rrdb_prepare')
        object.__setattr__(self, '_sig', sig)
        object.__setattr__(self, '_keep', keep)

    def __del__(self):
        try:
            if getattr(self, '_net', None) is not None:
                _lib.load().cfb_rrdb_destroy(self._net)
        except Exception:
            pass

    def forward(self, x):
        """x [B, num_in_ch, H, W] fp32 CUDA -> [B, num_out_ch, H*scale, W*scale] (rrdbnet_arch.py:103-119)."""
        if not (torch.is_tensor(x) and x.is_cuda):
            raise RuntimeError('RRDBNet.forward: codeformer_b200 runs on a CUDA device only; there is no CPU fallback')
        if x.dtype != torch.float32:
            raise RuntimeError(f'RRDBNet.forward: expected float32, got {x.dtype} (the B200 path computes in split-fp16 x3 with '
                               'fp32 accumulation; .half() models are not needed)')
        if x.dim() != 4 or x.shape[1] != self.num_in_ch:
            raise RuntimeError(f'RRDBNet.forward: expected [B,{self.num_in_ch},H,W], got {tuple(x.shape)}')
        us = 2 if self.scale == 2 else (4 if self.scale == 1 else 1)
        B, _, H, W = x.shape
        if H % us or W % us:
            raise AssertionError('pixel_unshuffle needs H and W divisible by the factor (arch_util.py:202)')
        lib = _lib.load()
        x = x.contiguous()
        dev = x.device
        with self._lock, torch.cuda.device(dev):
            self._prepare(dev)
            out = torch.empty((B, self.num_out_ch, H // us * 4, W // us * 4), dtype=torch.float32, device=dev)
            need = lib.cfb_rrdb_workspace_bytes(self._net, B, H, W)
            if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                object.__setattr__(self, '_ws', None)
                object.__setattr__(self, '_ws', torch.empty(int(need), dtype=torch.uint8, device=dev))
            _lib.check(lib.cfb_rrdb_forward(self._net, _lib.ptr(x), _lib.ptr(out), B, H, W, _lib.ptr(self._ws), self._ws.numel(),
                                            ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), 'cfb_rrdb_forward')
        return out


class RealESRGANer:
    """The reference's helper around the upsampling network (realesrgan_utils.py:14-250): ``enhance(img)`` takes an HWC
    uint8 / uint16 BGR (or gray, or BGRA) image and returns ``(upsampled image, mode)``.  ``model`` is any module mapping
    [1,3,h,w] -> [1,3,h*scale,w*scale] on ``device`` (``codeformer_b200.RRDBNet`` in production; the tests also pass CPU
    stand-ins to compare the tiling against the reference's)."""

    def __init__(self, scale, model_path=None, model=None, tile=0, tile_pad=10, pre_pad=10, half=False, device=None, gpu_id=None):
        self.scale, self.tile_size, self.tile_pad, self.pre_pad = scale, tile, tile_pad, pre_pad
        self.mod_scale = None
        self.half = False          # accepted for signature parity; the B200 path keeps fp32 semantics at tensor-core speed
        if device is None:
            device = torch.device('cuda', gpu_id if gpu_id is not None else torch.cuda.current_device())
        self.device = torch.device(device)
        if model_path is not None:                                  # realesrgan_utils.py:59-66
            loadnet = torch.load(model_path, map_location=torch.device('cpu'))
            model.load_state_dict(loadnet['params_ema' if 'params_ema' in loadnet else 'params'], strict=True)
        model.eval()
        self.model = model.to(self.device)

    def pre_process(self, img):
        """HWC float image -> [1,C,H,W] on the device, reflect pre-pad, reflect pad to the pixel-unshuffle multiple (:71-94)."""
        t = torch.from_numpy(np.ascontiguousarray(np.transpose(img, (2, 0, 1)))).float()
        self.img = t.unsqueeze(0).to(self.device)
        if self.pre_pad != 0:
            self.img = F.pad(self.img, (0, self.pre_pad, 0, self.pre_pad), 'reflect')
        self.mod_scale = 2 if self.scale == 2 else (4 if self.scale == 1 else None)
        if self.mod_scale is not None:
            _, _, h, w = self.img.shape
            self.mod_pad_h = (self.mod_scale - h % self.mod_scale) % self.mod_scale
            self.mod_pad_w = (self.mod_scale - w % self.mod_scale) % self.mod_scale
            self.img = F.pad(self.img, (0, self.mod_pad_w, 0, self.mod_pad_h), 'reflect')

    def process(self):
        self.output = self.model(self.img)

    def tile_plan(self, height, width):
        """Tile rectangles of tile_process (:100-175): for every tile (input rect with padding, output rect, crop of the
        model output).  Pure function of the sizes -- tested on the CPU against the reference's loop."""
        plan = []
        ts, tp, sc = self.tile_size, self.tile_pad, self.scale
        for y in range(math.ceil(height / ts)):
            for x in range(math.ceil(width / ts)):
                x0, x1 = x * ts, min(x * ts + ts, width)
                y0, y1 = y * ts, min(y * ts + ts, height)
                px0, px1 = max(x0 - tp, 0), min(x1 + tp, width)
                py0, py1 = max(y0 - tp, 0), min(y1 + tp, height)
                plan.append({'in': (py0, py1, px0, px1), 'out': (y0 * sc, y1 * sc, x0 * sc, x1 * sc),
                             'crop': ((y0 - py0) * sc, (y0 - py0) * sc + (y1 - y0) * sc, (x0 - px0) * sc, (x0 - px0) * sc + (x1 - x0) * sc)})
        return plan

    def tile_process(self):
        b, c, height, width = self.img.shape
        self.output = self.img.new_zeros((b, c, height * self.scale, width * self.scale))
        for t in self.tile_plan(height, width):
            py0, py1, px0, px1 = t['in']
            tile = self.model(self.img[:, :, py0:py1, px0:px1].contiguous())
            oy0, oy1, ox0, ox1 = t['out']
            cy0, cy1, cx0, cx1 = t['crop']
            self.output[:, :, oy0:oy1, ox0:ox1] = tile[:, :, cy0:cy1, cx0:cx1]

    def post_process(self):
        if self.mod_scale is not None:                              # :177-186
            _, _, h, w = self.output.shape
            self.output = self.output[:, :, 0:h - self.mod_pad_h * self.scale, 0:w - self.mod_pad_w * self.scale]
        if self.pre_pad != 0:
            _, _, h, w = self.output.shape
            self.output = self.output[:, :, 0:h - self.pre_pad * self.scale, 0:w - self.pre_pad * self.scale]
        return self.output

    def _run(self, img_rgb):
        self.pre_process(img_rgb)
        if self.tile_size > 0:
            self.tile_process()
        else:
            self.process()
        out = self.post_process().squeeze(0).float().cpu().clamp_(0, 1).numpy()
        return np.transpose(out[[2, 1, 0], :, :], (1, 2, 0))          # RGB CHW -> BGR HWC (:209-210)

    @torch.no_grad()
    def enhance(self, img, outscale=None, alpha_upsampler='realesrgan'):
        import cv2
        h_input, w_input = img.shape[0:2]
        img = img.astype(np.float32)
        max_range = 65535 if np.max(img) > 256 else 255              # :193-199
        img = img / max_range
        alpha = None
        if img.ndim == 2:
            img_mode, img = 'L', cv2.cvtColor(img, cv2.COLOR_GRAY2RGB)
        elif img.shape[2] == 4:
            img_mode, alpha = 'RGBA', img[:, :, 3]
            img = cv2.cvtColor(img[:, :, 0:3], cv2.COLOR_BGR2RGB)
            if alpha_upsampler == 'realesrgan':
                alpha = cv2.cvtColor(alpha, cv2.COLOR_GRAY2RGB)
        else:
            img_mode, img = 'RGB', cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        output_img = self._run(img)
        if img_mode == 'L':
            output_img = cv2.cvtColor(output_img, cv2.COLOR_BGR2GRAY)
        if img_mode == 'RGBA':                                        # :218-236
            if alpha_upsampler == 'realesrgan':
                output_alpha = cv2.cvtColor(self._run(alpha), cv2.COLOR_BGR2GRAY)
            else:
                h, w = alpha.shape[0:2]
                output_alpha = cv2.resize(alpha, (w * self.scale, h * self.scale), interpolation=cv2.INTER_LINEAR)
            output_img = cv2.cvtColor(output_img, cv2.COLOR_BGR2BGRA)
            output_img[:, :, 3] = output_alpha
        if max_range == 65535:
            output = (output_img * 65535.0).round().astype(np.uint16)
        else:
            output = (output_img * 255.0).round().astype(np.uint8)
        if outscale is not None and outscale != float(self.scale):
            output = cv2.resize(output, (int(w_input * outscale), int(h_input * outscale)), interpolation=cv2.INTER_LANCZOS4)
        return output, img_mode
