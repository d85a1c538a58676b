# -*- coding: utf-8 -*-
"""Host-side mirror of tasmania's backend / stencil registry for the ``b200`` backend.

tasmania selects a backend through four registries keyed by ``(function, backend, stencil)``
(src/tasmania/framework/protocol.py:L49-L59): allocators (``zeros/ones/empty/as_storage``,
src/tasmania/framework/allocators.py:L40-L181), ``stencil_definition``,
``stencil_compiler`` and their subroutine twins (src/tasmania/framework/stencil.py:L45-L204),
all reached through ``StencilFactory`` (stencil.py:L206-L469).  The reference itself cannot be
imported offline (gt4py / sympl / pint are absent), so this module restates
that *interface* -- same names, same argument meaning, same ``FactoryRegistryError`` when a
stencil is unknown -- for the one backend we ship; ``tasmania_b200.plugin`` performs the real
registration into tasmania when it is importable.

There is exactly one backend here and no dispatch: asking for anything but ``"b200"`` raises.
"""
from __future__ import annotations

import functools
import inspect
from dataclasses import dataclass, field
from typing import Any, Callable, Optional, Sequence

import numpy as np

from tasmania_b200 import storage

BACKEND = "b200"


class FactoryRegistryError(Exception):
    """Same role as src/tasmania/utils/exceptions.py:L41."""


@dataclass
class BackendOptions:
    """The fields of src/tasmania/framework/options.py:L48-L70 that a b200 stencil can see.
    ``externals`` carries the compile-time switches (flux scheme, ``moist``, physical
    constants, ...); the compiler snapshots it because components sharing one instance
    overwrite it right before each ``compile_stencil``."""

    dtypes: dict = field(default_factory=dict)
    exec_info: Optional[dict] = None
    externals: dict = field(default_factory=dict)
    rebuild: bool = False
    validate_args: bool = False
    verbose: bool = True


@dataclass
class StorageOptions:
    """src/tasmania/framework/options.py:L73-L81 (+ the device the storage lives on)."""

    dtype: Any = np.float64
    device: Optional[str] = None


# ------------------------------------------------------------------ registries
_STENCILS: dict[str, Callable] = {}
_SUBROUTINES: dict[str, Callable] = {}


def stencil_definition(stencil: str | Sequence[str]) -> Callable:
    """Register a b200 stencil definition (cf. ``StencilDefinition.register``,
    src/tasmania/framework/stencil.py:L112-L130).  The definition is called as
    ``definition(externals, **stencil_kwargs)``."""
    names = (stencil,) if isinstance(stencil, str) else tuple(stencil)

    def deco(fn):
        for n in names:
            _STENCILS[n] = fn
        fn.__tasmania_b200__ = {"function": "stencil_definition", "backend": BACKEND, "stencil": names}
        return fn

    return deco


def subroutine_definition(stencil: str | Sequence[str]) -> Callable:
    names = (stencil,) if isinstance(stencil, str) else tuple(stencil)

    def deco(obj):
        for n in names:
            _SUBROUTINES[n] = obj
        return obj

    return deco


def registered_stencils():
    return sorted(_STENCILS)


def _check_backend(backend):
    if backend not in (None, BACKEND):
        raise FactoryRegistryError(
            f"No compiler registered for the backend '{backend}': tasmania_b200 ships the "
            f"'{BACKEND}' backend only (no multi-backend dispatch, no CPU fallback)."
        )


def get_stencil_definition(stencil: str, backend: Optional[str] = None) -> Callable:
    _check_backend(backend)
    try:
        return _STENCILS[stencil]
    except KeyError:
        raise FactoryRegistryError(
            f"No definition of the stencil '{stencil}' found for the backend '{BACKEND}'."
        ) from None


def get_subroutine_definition(stencil: str, backend: Optional[str] = None):
    _check_backend(backend)
    try:
        return _SUBROUTINES[stencil]
    except KeyError:
        raise FactoryRegistryError(
            f"No definition of the subroutine '{stencil}' found for the backend '{BACKEND}'."
        ) from None


def compiler_b200(definition: Callable, *, backend_options: Optional[BackendOptions] = None) -> Callable:
    """The b200 ``stencil_compiler`` (cf. ``compiler_numpy`` + ``wrap``,
    src/tasmania/framework/subclasses/stencil_compilers.py:L49-L99): snapshot the externals,
    swallow the keyword arguments the definition does not name (``exec_info``,
    ``validate_args``, unused optional fields)."""
    bo = backend_options or BackendOptions()
    externals = dict(bo.externals or {})
    sig = inspect.signature(definition)
    accepted = {
        name for name, p in sig.parameters.items()
        if p.kind in (p.POSITIONAL_OR_KEYWORD, p.KEYWORD_ONLY)
    }
    has_var_kw = any(p.kind == p.VAR_KEYWORD for p in sig.parameters.values())

    @functools.wraps(definition)
    def stencil(**kwargs):
        if not has_var_kw:
            kwargs = {k: v for k, v in kwargs.items() if k in accepted}
        return definition(externals, **kwargs)

    stencil.externals = externals
    return stencil


def subroutine_compiler_b200(definition, *, backend_options: Optional[BackendOptions] = None):
    """The b200 ``subroutine_compiler``: a b200 subroutine is a scheme descriptor read by the fused
    kernels' marshalling code, never a callable to inline -- nothing to compile."""
    return definition


def compile_stencil(stencil: str, backend: Optional[str] = None, *,
                    backend_options: Optional[BackendOptions] = None) -> Callable:
    return compiler_b200(get_stencil_definition(stencil, backend), backend_options=backend_options)


class StencilFactory:
    """Mirror of src/tasmania/framework/stencil.py:L206-L469 for the b200 backend."""

    def __init__(self, backend: Optional[str] = None,
                 backend_options: Optional[BackendOptions] = None,
                 storage_options: Optional[StorageOptions] = None) -> None:
        _check_backend(backend)
        self._backend = BACKEND
        self._backend_options = backend_options or BackendOptions()
        self._storage_options = storage_options or StorageOptions()

    @property
    def backend(self) -> str:
        return self._backend

    @property
    def backend_options(self) -> BackendOptions:
        return self._backend_options

    @property
    def storage_options(self) -> StorageOptions:
        return self._storage_options

    # class-scoped definitions: several classes define the same stencil name ("saturation",
    # "diffusion", ...) and the per-instance registry tells them apart
    # (framework/stencil.py:L460-L469); here a class maps name -> registered b200 name
    class_stencils: dict = {}

    def compile_stencil(self, stencil: str, backend: Optional[str] = None, *,
                        backend_options: Optional[BackendOptions] = None) -> Callable:
        return compiler_b200(self.get_stencil_definition(stencil, backend),
                             backend_options=backend_options or self.backend_options)

    def get_stencil_definition(self, stencil, backend=None):
        return get_stencil_definition(self.class_stencils.get(stencil, stencil), backend)

    def get_subroutine_definition(self, stencil, backend=None):
        return get_subroutine_definition(stencil, backend)

    # framework/stencil.py:L286-L297, L356-L427: one compiler per kind for the one backend
    def get_stencil_compiler(self, backend=None, stencil=None):
        _check_backend(backend)
        return compiler_b200

    def get_subroutine_compiler(self, backend=None, stencil=None):
        _check_backend(backend)
        return subroutine_compiler_b200

    def compile_subroutine(self, stencil: str, backend: Optional[str] = None, *,
                           backend_options: Optional[BackendOptions] = None):
        return subroutine_compiler_b200(self.get_subroutine_definition(stencil, backend),
                                        backend_options=backend_options or self.backend_options)

    def _so(self, storage_options):
        return storage_options or self.storage_options

    def zeros(self, backend=None, *, shape, storage_options=None):
        _check_backend(backend)
        so = self._so(storage_options)
        return storage.zeros(shape, dtype=so.dtype, device=so.device)

    def ones(self, backend=None, *, shape, storage_options=None):
        _check_backend(backend)
        so = self._so(storage_options)
        return storage.ones(shape, dtype=so.dtype, device=so.device)

    def empty(self, backend=None, *, shape, storage_options=None):
        _check_backend(backend)
        so = self._so(storage_options)
        return storage.empty(shape, dtype=so.dtype, device=so.device)

    def as_storage(self, backend=None, *, data, storage_options=None):
        if backend == "numpy":
            return storage.to_numpy(data)
        _check_backend(backend)
        so = self._so(storage_options)
        return storage.as_storage(data, device=so.device)


class GridComponent:
    """Shape helpers of the reference's components over a grid
    (src/tasmania/framework/base_components.py:L55-L135): the extent of a named field on the grid --
    one more point along an axis it is staggered on, one level for a surface field, nz + 1 on the
    interface levels -- and the storage shape that holds it."""

    @property
    def grid(self):
        return self._grid

    @grid.setter
    def grid(self, value):
        self._grid = value

    def get_field_grid_shape(self, name: str):
        g = self.grid
        stag_x = "at_u_locations" in name or "at_uv_locations" in name
        stag_y = "at_v_locations" in name or "at_uv_locations" in name
        nk = 1 if "at_surface_level" in name else g.nz + 1 if "on_interface_levels" in name else g.nz
        return g.nx + int(stag_x), g.ny + int(stag_y), nk

    def get_field_storage_shape(self, name: str, default_storage_shape):
        return self.get_shape(default_storage_shape, min_shape=self.get_field_grid_shape(name))

    def get_storage_shape(self, shape, min_shape=None, max_shape=None):
        g = self.grid
        return self.get_shape(shape, min_shape or (g.nx, g.ny, g.nz), max_shape)

    @staticmethod
    def get_shape(in_shape, min_shape, max_shape=None):
        """``in_shape`` (or ``min_shape`` when none is given) clamped from below by ``min_shape`` and,
        if given, from above by ``max_shape``."""
        out = in_shape or min_shape
        if max_shape is None:
            return [max(a, lo) for a, lo in zip(out, min_shape)]
        return [lo if a < lo else (hi if a > hi else a) for a, lo, hi in zip(out, min_shape, max_shape)]

