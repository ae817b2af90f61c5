"""Registry with the interface of /root/reference/basicsr/utils/registry.py:4-82.

The reference looks its networks up with ``ARCH_REGISTRY.get('CodeFormer')`` /
``.get('VQAutoEncoder')`` (inference_codeformer.py:135, scripts/inference_vqgan.py:31); registering a
second class under an existing name asserts (registry.py:39).  ``install()`` therefore either
*replaces* the entries in the reference's own ARCH_REGISTRY (when ``basicsr`` is importable) or the
caller uses this module's ARCH_REGISTRY, which has the same methods.
"""


class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj):
        assert name not in self._obj_map, (f"An object named '{name}' was already registered "
                                           f"in '{self._name}' registry!")
        self._obj_map[name] = obj

    def register(self, obj=None):
        if obj is None:
            def deco(func_or_class):
                self._do_register(func_or_class.__name__, func_or_class)
                return func_or_class
            return deco
        self._do_register(obj.__name__, obj)

    def get(self, name):
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return ret

    def __contains__(self, name):
        return name in self._obj_map

    def __iter__(self):
        return iter(self._obj_map.items())

    def keys(self):
        return self._obj_map.keys()


ARCH_REGISTRY = Registry('arch')


def install(registry=None):
    """Make ``registry.get('CodeFormer' | 'VQAutoEncoder')`` return the B200 modules.

    With ``registry=None`` the reference's own ``basicsr.utils.registry.ARCH_REGISTRY`` is patched if
    ``basicsr`` is importable; the existing entries are replaced (not added beside: registry.py:39)."""
    from .arch import CodeFormer, VQAutoEncoder
    if registry is None:
        from basicsr.utils.registry import ARCH_REGISTRY as registry  # type: ignore
    for cls in (CodeFormer, VQAutoEncoder):
        registry._obj_map[cls.__name__] = cls
    return registry
