"""codeformer_b200 -- B200-native (sm_100a) execution of CodeFormer's core forward pass.

Public surface = the reference's plugin API for this path:
    ARCH_REGISTRY.get('CodeFormer') / .get('VQAutoEncoder')     basicsr/utils/registry.py:62
    CodeFormer(...).forward(x, w, detach_16, code_only, adain)  basicsr/archs/codeformer_arch.py:223
    VQAutoEncoder(...).forward(x)                               basicsr/archs/vqgan_arch.py:385
"""
from .registry import ARCH_REGISTRY, install          # noqa: F401
from .arch import CodeFormer, VQAutoEncoder, VectorQuantizer   # noqa: F401
from .upsampler import RRDBNet, RealESRGANer                   # noqa: F401
from .parsing import ParseNet, face_parse_mask, init_parsing_model   # noqa: F401


def check_async_status():
    """Raise ``RuntimeError`` if a kernel of an earlier (asynchronous) forward on the current device reported a failure --
    a tensor-core pipeline time-out or an activation outside the fp16 operand range.  Call after synchronising the stream;
    the next forward and the ``restore_faces`` / ``forward_host`` front-ends check by themselves (include/cfb200.h)."""
    from . import _lib
    _lib.check(_lib.load().cfb_check_async_status(), 'check_async_status')


__all__ = ['ARCH_REGISTRY', 'install', 'CodeFormer', 'VQAutoEncoder', 'VectorQuantizer', 'RRDBNet', 'RealESRGANer', 'ParseNet', 'face_parse_mask', 'init_parsing_model',
           'check_async_status']
