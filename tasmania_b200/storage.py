# -*- coding: utf-8 -*-
"""Device storages of the ``b200`` backend.

tasmania treats storages as numpy-like duck arrays: ``.shape``, ``.dtype``, slicing that
returns stride-correct views, ``a[...] = scalar | ndarray | storage`` with broadcasting,
``deepcopy`` and conversion to numpy (e.g.
src/tasmania/domain/subclasses/horizontal_boundaries/relaxed.py:L236-L247,
src/tasmania/isentropic/dynamics/diagnostics.py:L118-L121).  ``B200Array`` provides exactly
that on top of a ``torch`` CUDA tensor, which is used for *allocation only*; all numerics
go through the CUDA kernels behind the C ABI.

Layout (ours to choose because we own the allocator): logical shape ``(ni, nj, nk)`` with
**i fastest**, rows padded to a multiple of 16 doubles (128 bytes), i.e. element strides
``(1, ni_pad, ni_pad * nj)``.  Horizontal stencils then read coalesced 128-byte rows and the
vertical scans run one thread per column with i along the warp.  (numpy's C order makes k
fastest -- fine for numpy, wrong for CUDA.)
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

ROW_ALIGN = 16  # doubles -> 128 bytes

# Host-side tests of the storage wrapper / the tasmania plugin (no GPU in the build container)
# set this to "cpu"; the product never does, and kernels refuse host memory (lib.as_field).
DEFAULT_DEVICE_OVERRIDE = None


def _round_up(n: int, m: int) -> int:
    return ((n + m - 1) // m) * m


def default_device() -> torch.device:
    if DEFAULT_DEVICE_OVERRIDE is not None:
        return torch.device(DEFAULT_DEVICE_OVERRIDE)
    if not torch.cuda.is_available():
        raise RuntimeError(
            "the b200 backend needs a CUDA device (there is no CPU fallback); pass "
            "device='cpu' explicitly only to exercise the storage wrapper in host-side tests"
        )
    return torch.device("cuda", torch.cuda.current_device())


def _torch_dtype(dtype) -> torch.dtype:
    dt = np.dtype(dtype)
    table = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
             np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
             np.dtype(np.bool_): torch.bool}
    if dt not in table:
        raise TypeError(f"unsupported storage dtype {dt}")
    return table[dt]


class B200Array:
    """A strided fp64 view of device memory with numpy-like slicing semantics."""

    __slots__ = ("t",)
    __array_priority__ = 1000

    def __init__(self, tensor: torch.Tensor):
        self.t = tensor

    # ---- numpy-like metadata
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def ndim(self):
        return self.t.dim()

    @property
    def size(self):
        return self.t.numel()

    @property
    def dtype(self):
        return np.dtype(str(self.t.dtype).replace("torch.", ""))

    @property
    def strides(self):
        es = self.t.element_size()
        return tuple(s * es for s in self.t.stride())

    @property
    def device(self):
        return self.t.device

    @property
    def __cuda_array_interface__(self):
        if not self.t.is_cuda:
            raise AttributeError("__cuda_array_interface__ (storage lives on the host)")
        return {
            "shape": self.shape,
            "typestr": self.dtype.str,
            "data": (self.t.data_ptr(), False),
            "strides": self.strides,
            "version": 3,
        }

    # ---- conversions
    def to_numpy(self) -> np.ndarray:
        return self.t.detach().cpu().numpy().copy()

    def __array__(self, dtype=None, copy=None):
        out = self.to_numpy()
        return out if dtype is None else out.astype(dtype)

    def item(self):
        return self.t.item()

    def __float__(self):
        return float(self.t.item())

    def __len__(self):
        return self.t.shape[0]

    def __repr__(self):
        return f"B200Array(shape={self.shape}, strides={self.t.stride()}, device={self.t.device})"

    # ---- indexing
    @staticmethod
    def _norm_index(idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        out = []
        for it in idx:
            if it is np.newaxis:
                out.append(None)
            elif isinstance(it, (np.integer,)):
                out.append(int(it))
            elif isinstance(it, B200Array):
                out.append(it.t)
            elif isinstance(it, np.ndarray):
                out.append(torch.as_tensor(it))
            else:
                out.append(it)
        return tuple(out)

    def __getitem__(self, idx):
        res = self.t[self._norm_index(idx)]
        if res.dim() == 0:
            return res.item()
        return B200Array(res)

    def _coerce(self, value) -> torch.Tensor | float:
        if isinstance(value, B200Array):
            return value.t
        if isinstance(value, torch.Tensor):
            return value.to(self.t.device)
        if isinstance(value, np.ndarray):
            return torch.as_tensor(np.ascontiguousarray(value), dtype=self.t.dtype).to(self.t.device)
        if isinstance(value, (list, tuple)):
            return torch.as_tensor(np.asarray(value), dtype=self.t.dtype).to(self.t.device)
        return value  # python / numpy scalar

    def __setitem__(self, idx, value):
        self.t[self._norm_index(idx)] = self._coerce(value)

    # ---- copies
    def copy(self) -> "B200Array":
        out = empty(self.shape, dtype=self.dtype, device=self.t.device) if self.ndim == 3 else \
            B200Array(torch.empty_like(self.t))
        out.t.copy_(self.t)
        return out

    def __deepcopy__(self, memo):
        return self.copy()

    def __copy__(self):
        return B200Array(self.t)

    # ---- a minimum of arithmetic for host-side set-up code (never on the hot path)
    def _bin(self, other, op):
        return B200Array(op(self.t, self._coerce(other)))

    def __add__(self, o):
        return self._bin(o, torch.add)

    __radd__ = __add__

    def __sub__(self, o):
        return self._bin(o, torch.sub)

    def __rsub__(self, o):
        return B200Array(torch.sub(torch.as_tensor(self._coerce(o), device=self.t.device), self.t))

    def __mul__(self, o):
        return self._bin(o, torch.mul)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._bin(o, torch.div)

    def __neg__(self):
        return B200Array(-self.t)

    def __iadd__(self, o):
        self.t += self._coerce(o)
        return self

    def __isub__(self, o):
        self.t -= self._coerce(o)
        return self

    def __imul__(self, o):
        self.t *= self._coerce(o)
        return self

    def max(self):
        return self.t.max().item()

    def min(self):
        return self.t.min().item()


# ------------------------------------------------------------------ allocators
def _allocate(shape: Sequence[int], dtype, device, fill) -> B200Array:
    dev = torch.device(device) if device is not None else default_device()
    tdt = _torch_dtype(dtype)
    shape = tuple(int(n) for n in shape)
    if len(shape) != 3:
        base = torch.empty(shape, dtype=tdt, device=dev)
        if fill is not None:
            base.fill_(fill)
        return B200Array(base)
    ni, nj, nk = shape
    ni_pad = _round_up(max(ni, 1), ROW_ALIGN)
    base = torch.empty((max(nk, 1), max(nj, 1), ni_pad), dtype=tdt, device=dev)
    if fill is not None:
        base.fill_(fill)
    return B200Array(base.permute(2, 1, 0)[:ni, :nj, :nk])


def empty(shape, *, dtype=np.float64, device=None) -> B200Array:
    return _allocate(shape, dtype, device, None)


def zeros(shape, *, dtype=np.float64, device=None) -> B200Array:
    return _allocate(shape, dtype, device, 0)


def ones(shape, *, dtype=np.float64, device=None) -> B200Array:
    return _allocate(shape, dtype, device, 1)


def as_storage(data, *, dtype=None, device=None) -> B200Array:
    """Host or device data -> b200 storage (H2D copy into the i-fastest padded layout)."""
    if isinstance(data, B200Array):
        if device is None or torch.device(device) == data.t.device:
            return data
        out = empty(data.shape, dtype=data.dtype, device=device)
        out.t.copy_(data.t)
        return out
    if isinstance(data, torch.Tensor):
        arr = data
        out = _allocate(tuple(arr.shape), dtype or str(arr.dtype).replace("torch.", ""), device, None)
        out.t.copy_(arr)
        return out
    arr = np.asarray(data)
    dt = dtype or (arr.dtype if arr.dtype.kind == "f" else np.float64)
    out = _allocate(arr.shape, dt, device, None)
    src = torch.as_tensor(np.ascontiguousarray(arr).astype(dt, copy=False))
    if out.t.is_cuda:
        src = src.pin_memory() if src.numel() > (1 << 16) else src
        out.t.copy_(src, non_blocking=False)
    else:
        out.t.copy_(src)
    return out


def to_numpy(x) -> np.ndarray:
    """Device -> host (D2H copy).  The b200 overload of
    src/tasmania/framework/generic_functions.py:L35-L36."""
    if isinstance(x, B200Array):
        return x.to_numpy()
    return np.asarray(x)


def stage_scratch(shape, count, device=None):
    """``count`` hand-off fields of the fused RK stage for a host-side object that issues the fused
    calls: ``(context, fields)``.  On the current CUDA device they come from a library context
    (``tb200_ctx_create / scratch / destroy``; the caller keeps ``context``
    alive and drops it with itself); storages of another device, the CPU test double of the
    library (``DEFAULT_DEVICE_OVERRIDE``) and ``TB200_CTX_SCRATCH=0`` keep plain storages
    (``context`` is None).  Call it outside any CUDA-graph capture."""
    import os

    from tasmania_b200 import lib

    use_ctx = DEFAULT_DEVICE_OVERRIDE is None and os.environ.get("TB200_CTX_SCRATCH", "1") != "0"
    if use_ctx and device is not None:
        d = torch.device(device)
        # the context allocates on the CURRENT device
        use_ctx = d.type == "cuda" and (d.index is None or d.index == torch.cuda.current_device())
    if use_ctx:
        ctx = lib.Context()
        return ctx, ctx.scratch(tuple(int(n) for n in shape), count)
    return None, tuple(zeros(tuple(shape), device=device) for _ in range(count))

