"""Module objects whose attributes resolve lazily to objects that raise when called (plotting is out of scope)."""
import sys
import types


class _Inert:
    def __init__(self, name):
        self._name = name

    def __getattr__(self, k):
        return _Inert(self._name + "." + k)

    def __call__(self, *a, **kw):
        raise RuntimeError("%s is a stub (compat/optional_stubs): plotting / tables are outside the GPU hot path" % self._name)


def make(name, submodules=()):
    m = types.ModuleType(name)
    m.__getattr__ = lambda k, _n=name: _Inert(_n + "." + k)  # PEP 562
    for s in submodules:
        sm = types.ModuleType(name + "." + s)
        sm.__getattr__ = lambda k, _n=name + "." + s: _Inert(_n + "." + k)
        sys.modules[name + "." + s] = sm
        setattr(m, s, sm)
    return m
