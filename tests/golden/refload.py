# -*- coding: utf-8 -*-
"""Import the reference's own numpy stencils *in place* from /root/reference.

TEST INFRASTRUCTURE ONLY.  This module is used exclusively by
``tests/golden/generate_golden.py`` (fixture generation, run in the build
container where ``/root/reference`` is mounted) and by the optional
``tests/test_plugin_reference.py``.  Nothing in the product or in the ``-m gpu``
tests / ``bench.py`` / ``smoke()`` imports it: ``/root/reference`` does not exist
on the GPU box.

The reference cannot be imported as shipped (``import tasmania`` pulls
``gt4py``, the private ``sympl`` fork, ``pint``, ``xarray`` ... none installable
offline).  Its numpy backend however is plain numpy.  We therefore

* register every ``tasmania.*`` package as a *namespace-only* module whose
  ``__path__`` points into ``/root/reference/src/tasmania`` (so no
  ``__init__.py`` is executed and only the sub-modules we ask for are loaded),
* provide tiny stand-ins for ``gt4py.cartesian.gtscript`` (decorators become the
  identity) and for the handful of ``sympl`` names the numerics touch
  (``DataArray`` with ``to_units``, ``AbstractFactory`` with a name-keyed
  ``factory()``, a no-op ``Timer``).

No reference source is copied: the reference code is executed where it lies.
"""
from __future__ import annotations

import abc
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("TASMANIA_REFERENCE", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "tasmania"))


# --------------------------------------------------------------------------- stubs
class _AnyMeta(abc.ABCMeta):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _Anything(metaclass=_AnyMeta):
    """Subclassable, callable, subscriptable placeholder."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __class_getitem__(cls, item):
        return cls

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubModule(types.ModuleType):
    """Module whose unknown attributes resolve to fresh placeholder classes."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (_Anything,), {})
        setattr(self, name, obj)
        return obj


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ("gt4py", "sympl", "pint", "xarray", "netCDF4")

    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in self.PREFIXES:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _StubModule(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        _populate_stub(module)


_UNIT_SCALE = {
    # (from, to): factor -- only what the numerics' constructors ask for
    ("km", "m"): 1.0e3,
    ("m", "km"): 1.0e-3,
    ("hPa", "Pa"): 1.0e2,
    ("g g^-1", "kg kg^-1"): 1.0,
    ("kg kg^-1", "g g^-1"): 1.0,
    ("g kg^-1", "g g^-1"): 1.0e-3,
    ("g kg^-1", "kg kg^-1"): 1.0e-3,
    ("mm h^-1", "mm hr^-1"): 1.0,
    ("mm hr^-1", "mm h^-1"): 1.0,
}


class DataArray:
    """Very small stand-in for sympl.DataArray (numpy payload + dims + units)."""

    def __init__(self, data, coords=None, dims=None, name=None, attrs=None):
        if isinstance(data, DataArray):
            data = data.data
        self.data = data if hasattr(data, "shape") else np.asarray(data, dtype=float)
        self.coords = coords
        if isinstance(dims, str):
            dims = (dims,)
        self.dims = tuple(dims) if dims is not None else tuple(
            f"dim_{i}" for i in range(np.ndim(self.data))
        )
        self.name = name
        self.attrs = dict(attrs or {})
        # xarray-like coords: {dim: DataArray}
        if isinstance(coords, (list, tuple)) and len(coords) == len(self.dims):
            self.coords = {
                d: (c if isinstance(c, DataArray) else DataArray(np.asarray(c), None, (d,)))
                for d, c in zip(self.dims, coords)
            }
        elif isinstance(coords, dict):
            self.coords = {
                d: (c if isinstance(c, DataArray) else DataArray(np.asarray(c), None, (d,)))
                for d, c in coords.items()
            }

    @property
    def values(self):
        return self.data

    @property
    def shape(self):
        return self.data.shape

    def item(self):
        return self.data.item()

    def to_units(self, units):
        have = self.attrs.get("units", units)
        if have == units:
            return self
        key = (have, units)
        if key not in _UNIT_SCALE:
            raise ValueError(f"refload stub: no conversion {have!r} -> {units!r}")
        return DataArray(self.data * _UNIT_SCALE[key], self.coords, self.dims, self.name,
                         {**self.attrs, "units": units})

    def __getitem__(self, idx):
        return DataArray(self.data[idx], None, None, self.name, self.attrs)

    def __setitem__(self, idx, value):
        self.data[idx] = value.data if isinstance(value, DataArray) else value

    def __getattr__(self, name):
        # dtype, ndim, size, min, max ... are answered by the payload
        if name.startswith("__") or name == "data":
            raise AttributeError(name)
        return getattr(self.data, name)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.data, dtype=dtype)

    def __bool__(self):
        return bool(self.data.all()) if self.data.ndim == 0 else True

    def _wrap(self, out):
        return DataArray(out, self.coords, self.dims if np.ndim(out) == len(self.dims) else None,
                         self.name, self.attrs)

    @staticmethod
    def _raw(x):
        return x.data if isinstance(x, DataArray) else x

    def __add__(self, o):
        return self._wrap(self.data + self._raw(o))

    __radd__ = __add__

    def __sub__(self, o):
        return self._wrap(self.data - self._raw(o))

    def __rsub__(self, o):
        return self._wrap(self._raw(o) - self.data)

    def __mul__(self, o):
        return self._wrap(self.data * self._raw(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._wrap(self.data / self._raw(o))

    def __rtruediv__(self, o):
        return self._wrap(self._raw(o) / self.data)

    def __neg__(self):
        return self._wrap(-self.data)

    def __deepcopy__(self, memo):
        import copy as _copy

        return DataArray(_copy.deepcopy(self.data, memo), _copy.deepcopy(self.coords, memo),
                         self.dims, self.name, dict(self.attrs))


class _FactoryMeta(type):
    pass


class AbstractFactory:
    """Name-keyed factory, as the private sympl fork's ``sympl._core.factory`` provides."""

    name = None

    @classmethod
    def _all_subclasses(cls):
        out = []
        for sub in cls.__subclasses__():
            out.append(sub)
            out.extend(sub._all_subclasses())
        return out

    @classmethod
    def factory(cls, name, *args, **kwargs):
        for sub in cls._all_subclasses():
            if getattr(sub, "name", None) == name:
                return sub(*args, **kwargs)
        raise KeyError(f"refload stub: no subclass of {cls.__name__} named {name!r}")


class Timer:
    @classmethod
    def start(cls, label=None, *a, **k):
        pass

    @classmethod
    def stop(cls, label=None, *a, **k):
        pass

    @classmethod
    def reset(cls, *a, **k):
        pass


def _identity_decorator(*dargs, **dkwargs):
    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]

    def deco(fn):
        return fn

    return deco


class _Field:
    def __class_getitem__(cls, item):
        return cls

    def __getitem__(self, item):
        return self


def _populate_stub(module):
    name = module.__name__
    if name == "gt4py.cartesian.gtscript":
        module.function = _identity_decorator
        module.stencil = _identity_decorator
        module.lazy_stencil = _identity_decorator
        module.Field = _Field
        module.Sequence = _Field
        for sym in ("PARALLEL", "FORWARD", "BACKWARD", "computation", "interval", "IJ", "IJK", "K", "I", "J"):
            setattr(module, sym, _Anything())
    elif name == "gt4py.cartesian":
        module.gtscript = importlib.import_module("gt4py.cartesian.gtscript")
    elif name == "sympl":
        module.DataArray = DataArray

        def get_constant(name, units=None):
            # no sympl constant table here -> tasmania falls back to its own defaults
            raise KeyError(name)

        def set_constant(name, value, units=None):
            pass

        module.get_constant = get_constant
        module.set_constant = set_constant
    elif name == "sympl._core.data_array":
        module.DataArray = DataArray
    elif name == "sympl._core.factory":
        module.AbstractFactory = AbstractFactory
    elif name == "sympl._core.time":
        module.Timer = Timer
        module.FakeTimer = Timer


_installed = False


def install():
    """Idempotently make ``import tasmania.<sub>.<module>`` resolve into the reference tree."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_SRC}")
    sys.meta_path.insert(0, _StubFinder())

    base = os.path.join(REF_SRC, "tasmania")
    # namespace-only packages: set __path__, never execute __init__.py
    for dirpath, dirnames, filenames in os.walk(base):
        dirnames[:] = [d for d in dirnames if not d.startswith("__")]
        rel = os.path.relpath(dirpath, base)
        modname = "tasmania" if rel == "." else "tasmania." + rel.replace(os.sep, ".")
        mod = types.ModuleType(modname)
        mod.__path__ = [dirpath]
        mod.__package__ = modname
        sys.modules[modname] = mod
    # wire children as attributes so ``from tasmania.framework import protocol`` works
    for modname, mod in list(sys.modules.items()):
        if modname.startswith("tasmania.") and "." in modname:
            parent, _, child = modname.rpartition(".")
            if parent in sys.modules:
                setattr(sys.modules[parent], child, mod)
    _installed = True


def load(modname):
    install()
    return importlib.import_module(modname)


def numpy_stencil(definition, externals=None):
    """Bind ``externals`` into the definition's globals -- exactly what the reference's
    ``compiler_numpy`` does (src/tasmania/framework/subclasses/stencil_compilers.py:L91-L99)."""
    if externals:
        definition.__globals__.update(externals)
    return definition


_FRAMEWORK_MODULES = (
    "tasmania.framework.subclasses.allocators.as_storage",
    "tasmania.framework.subclasses.allocators.as_storage_numpy",
    "tasmania.framework.subclasses.allocators.empty",
    "tasmania.framework.subclasses.allocators.ones",
    "tasmania.framework.subclasses.allocators.zeros",
    "tasmania.framework.subclasses.stencil_compilers",
    "tasmania.framework.subclasses.subroutine_compilers",
    "tasmania.framework.subclasses.stencil_definitions.algorithms",
    "tasmania.framework.subclasses.stencil_definitions.cla",
    "tasmania.framework.subclasses.stencil_definitions.copy",
    "tasmania.framework.subclasses.stencil_definitions.diffusion",
    "tasmania.framework.subclasses.stencil_definitions.math",
    "tasmania.framework.subclasses.subroutine_definitions.generics",
    "tasmania.framework.subclasses.subroutine_definitions.math",
    "tasmania.framework.subclasses.subroutine_definitions.laplacian",
    "tasmania.framework.subclasses.subroutine_definitions.cla",
)


def install_framework():
    """Load the reference's numpy allocators / compilers / global stencils, i.e. what
    ``tasmania/framework/__init__.py`` would register (minus gt4py/cupy/numba)."""
    install()
    loaded = []
    for name in _FRAMEWORK_MODULES:
        try:
            loaded.append(importlib.import_module(name))
        except ModuleNotFoundError:
            pass
    return loaded
