"""Name -> class lookup behind the reference's plugin API.

The reference finds its networks with ``ARCH_REGISTRY.get('CodeFormer')`` / ``.get('VQAutoEncoder')``
(/root/reference/inference_codeformer.py:135, scripts/inference_vqgan.py:31; the registry itself is
basicsr/utils/registry.py:4-82, whose ``register`` refuses a second class under an existing name, :39).
Two ways to drop the B200 modules in:

* ``install()`` swaps the two entries inside the reference's OWN ``ARCH_REGISTRY`` (when ``basicsr`` is
  importable) -- the caller's ``ARCH_REGISTRY.get(...)`` line stays as it is;
* without ``basicsr``, ``codeformer_b200.ARCH_REGISTRY`` below answers the same three calls the callers make
  (``register()`` as a decorator, ``get(name)``, ``name in registry``).
"""
from collections import OrderedDict


class ArchTable:
    """Insertion-ordered name -> class table with the reference registry's call surface."""

    def __init__(self, label):
        self.label = label
        self.table = OrderedDict()

    def add(self, cls, name=None):
        key = name or cls.__name__
        if key in self.table:
            raise AssertionError(f'{self.label}: {key!r} is already taken by {self.table[key]!r}')
        self.table[key] = cls
        return cls

    def register(self, cls=None):
        """``@REG.register()`` (decorator factory, the form the reference uses) or ``REG.register(cls)``."""
        return self.add if cls is None else self.add(cls)

    def get(self, name):
        try:
            return self.table[name]
        except KeyError:
            raise KeyError(f'{self.label}: nothing registered under {name!r} (have: {", ".join(self.table)})') from None

    def __contains__(self, name):
        return name in self.table

    def __iter__(self):
        return iter(self.table.items())

    def keys(self):
        return self.table.keys()


ARCH_REGISTRY = ArchTable('codeformer_b200 arch table')


def install(registry=None):
    """Make ``registry.get('CodeFormer' | 'VQAutoEncoder')`` return the B200 modules.

    ``registry=None`` patches the reference's ``basicsr.utils.registry.ARCH_REGISTRY`` (``basicsr`` must be importable).
    Existing entries are REPLACED -- adding beside them is impossible (registry.py:39 asserts on duplicates) -- by writing
    the mapping the reference registry keeps (its ``_obj_map`` dict) or, for any other mapping-like object, by item
    assignment."""
    from .arch import CodeFormer, VQAutoEncoder
    if registry is None:
        from basicsr.utils.registry import ARCH_REGISTRY as registry  # type: ignore
    store = getattr(registry, '_obj_map', None)
    if store is None:
        store = getattr(registry, 'table', registry)
    for cls in (CodeFormer, VQAutoEncoder):
        store[cls.__name__] = cls
    return registry
