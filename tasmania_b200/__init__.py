# -*- coding: utf-8 -*-
"""tasmania_b200 -- a B200-native (sm_100a, fp64) ``b200`` backend for the stencil hot path
of stubbiali/tasmania: hand-written CUDA kernels behind a C ABI (``include/tasmania_b200.h``),
a thin ctypes binding, device storages and the host-side mirror of tasmania's backend /
stencil registry.  There is no CPU fallback: computing without the CUDA library or without a
GPU raises.
"""
from tasmania_b200 import stencils as _stencils  # noqa: F401  (registers the definitions)
from tasmania_b200.framework import (  # noqa: F401
    BACKEND,
    BackendOptions,
    FactoryRegistryError,
    StencilFactory,
    StorageOptions,
    compile_stencil,
    get_stencil_definition,
    registered_stencils,
)
from tasmania_b200.lib import B200Error  # noqa: F401
from tasmania_b200.storage import B200Array, as_storage, empty, ones, to_numpy, zeros  # noqa: F401

__version__ = "0.1.0"
