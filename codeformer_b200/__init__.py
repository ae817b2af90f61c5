"""codeformer_b200 -- B200-native (sm_100a) execution of CodeFormer's core forward pass.

Public surface = the reference's plugin API for this path:
    ARCH_REGISTRY.get('CodeFormer') / .get('VQAutoEncoder')     basicsr/utils/registry.py:62
    CodeFormer(...).forward(x, w, detach_16, code_only, adain)  basicsr/archs/codeformer_arch.py:223
    VQAutoEncoder(...).forward(x)                               basicsr/archs/vqgan_arch.py:385
"""
from .registry import ARCH_REGISTRY, install          # noqa: F401
from .arch import CodeFormer, VQAutoEncoder, VectorQuantizer   # noqa: F401

__all__ = ['ARCH_REGISTRY', 'install', 'CodeFormer', 'VQAutoEncoder', 'VectorQuantizer']
