# -*- coding: utf-8 -*-
"""ctypes binding of ``libtasmania_b200.so`` (the C ABI declared in include/tasmania_b200.h).

There is no CPU fallback: if the library is missing, or a field handed to a kernel does not
live in device memory, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtasmania_b200.so")


class B200Error(RuntimeError):
    """A kernel launch was refused (bad arguments) or failed (CUDA error)."""


class Field(C.Structure):
    """``tb200_field``"""

    _fields_ = [("ptr", C.c_void_p), ("shape", C.c_int64 * 3), ("stride", C.c_int64 * 3)]


FieldP = C.POINTER(Field)
Int3 = C.c_int32 * 3
Double4 = C.c_double * 4


class StageCfg(C.Structure):
    """``tb200_isentropic_stage``"""

    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("nb", C.c_int32),
        ("flux_scheme", C.c_int32), ("damp", C.c_int32),
        ("dt", C.c_double), ("dt_full", C.c_double),
        ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double), ("eps", C.c_double),
        ("pt", C.c_double), ("theta_s", C.c_double),
        ("constants", C.c_double * 4),
        ("part", C.c_int32), ("rim", C.c_int32 * 4),
        ("derive_uv_in", C.c_int32), ("skip_uv_out", C.c_int32),
        ("s_tnd", FieldP), ("su_tnd", FieldP), ("sv_tnd", FieldP),
        ("periodic", C.c_int32),
    ]


class HaloSide(C.Structure):
    """``tb200_halo_side``"""

    _fields_ = [("remote_buffer", C.c_void_p), ("remote_counter", C.c_void_p),
                ("local_buffer", C.c_void_p), ("local_counter", C.c_void_p),
                ("slot_doubles", C.c_int64),
                ("send_origin", C.c_int32 * 2), ("recv_origin", C.c_int32 * 2), ("extent", C.c_int32 * 2)]


P2P_HANDLE_BYTES, P2P_CHANNEL_BYTES = 64, 32

FLUX_SCHEMES = {"upwind": 0, "centered": 1, "third_order_upwind": 2, "fifth_order_upwind": 3}
ELEMENTWISE_OPS = {
    "copy": 0, "copychange": 1, "abs": 2, "add": 3, "addsub": 4, "clip": 5, "fma": 6,
    "mul": 7, "scale": 8, "sub": 9, "sts_rk2_0": 10, "sts_rk3ws_0": 11, "iaddsub": 12,
    "iscale": 13,
}

_F, _I3, _D, _I, _V = FieldP, C.POINTER(C.c_int32), C.c_double, C.c_int, C.c_void_p
_FPP = C.POINTER(FieldP)

# every symbol include/tasmania_b200.h declares: name -> argtypes
SIGNATURES = {
    "tb200_elementwise": [_I, _F, _F, _F, _F, _D, _I3, _I3, _V],
    "tb200_fma_fields": [_I, _FPP, _FPP, _FPP, _D, _I3, _I3, _V],
    "tb200_relax": [_F, _F, _F, _F, _I3, _I3, _V],
    "tb200_relax_frame": [_I, _FPP, _FPP, _F, _I3, _I3, _V],
    "tb200_periodic_enforce": [_F, _I, _I, _I, _I, _I, _V],
    "tb200_set_outermost_layers": [_F, _F, _I, _I, _I, _V],
    "tb200_damping": [_F, _F, _F, _F, _F, _D, _I3, _I3, _V],
    "tb200_velocity": [_I, _F, _F, _F, _I, _I3, _I3, _V],
    "tb200_velocity_components": [_F] * 7 + [_I, _I, _I, _V],
    "tb200_momenta": [_F, _F, _F, _F, _F, _I, _I3, _I3, _V],
    "tb200_density": [_F, _F, _F, _I, _I3, _I3, _V],
    "tb200_mass_fraction": [_F, _F, _F, _I, _I3, _I3, _V],
    "tb200_diffusion": [_I, _F, _F, _F, _D, _D, _I, _I3, _I3, _V],
    "tb200_smoothing": [_I, _F, _F, _F, _I, _I3, _I3, _V],
    "tb200_hyperdiffusion": [_F, _F, _D, _I3, _I3, _V],
    "tb200_thomas": [_F, _F, _F, _F, _F, _I3, _I3, _V],
    "tb200_diffusion_1d": [_I, _I, _F, _F, _F, _D, _I, _I3, _I3, _V],
    "tb200_smoothing_1d": [_I, _I, _F, _F, _F, _I, _I3, _I3, _V],
    "tb200_step_forward_euler": [_I, _F, _F, _F, _F, _F, _F, _FPP, _FPP, _FPP, _FPP, _D, _D, _D,
                                 _I3, _I3, _V],
    "tb200_step_forward_euler_momentum": [_I] + [_F] * 14 + [_D, _D, _D, _D, _I3, _I3, _V],
    "tb200_montgomery": [_F, _F, _F, _D, _D, _D, C.POINTER(C.c_double), _I3, _I3, _V],
    "tb200_diagnostic_variables": [_F] * 7 + [_D, _D, C.POINTER(C.c_double), _I3, _I3, _V],
    "tb200_height": [_F] * 4 + [_D, _D, C.POINTER(C.c_double), _I3, _I3, _V],
    "tb200_density_and_temperature": [_F] * 6 + [_D, _I3, _I3, _V],
    "tb200_burgers_forward_euler": [_I] + [_F] * 8 + [_D, _D, _D, _I3, _I3, _V],
    "tb200_isentropic_stage_dry": [C.POINTER(StageCfg)] + [_F] * 25 + [_V],
    "tb200_isentropic_stage_moist": [C.POINTER(StageCfg)] + [_F] * 25 + [_FPP] * 4 + [_V],
    "tb200_stage_profile": [_I],
    "tb200_stage_lazy_velocities": [_I],
    "tb200_stage_profile_read": [C.POINTER(C.c_double)],
    "tb200_pack_box": [_F, C.c_void_p, _I3, _I3, _V],
    "tb200_unpack_box": [_F, C.c_void_p, _I3, _I3, _V],
    "tb200_kessler": [_F] * 11 + [_D] * 5 + [C.c_uint32, _I3, _I3, _V],
    "tb200_saturation_diagnostic": [_F] * 9 + [_D] * 5 + [C.c_uint32, _I3, _I3, _V],
    "tb200_saturation_prognostic": [_F] * 8 + [_D] * 5 + [C.c_uint32, _I3, _I3, _V],
    "tb200_fall_velocity": [_F] * 4 + [_I3, _I3, _V],
    "tb200_sedimentation": [_I] + [_F] * 5 + [_I, _I3, _I3, _V],
    "tb200_accumulated_precipitation": [_F] * 6 + [_D, _D, _I3, _I3, _V],
    "tb200_smagorinsky": [_F] * 5 + [_D, _D, _D, _I, _I, _I3, _I3, _V],
    "tb200_coriolis": [_F] * 4 + [_D, _I, _I, _I3, _I3, _V],
    "tb200_coriolis_step": [_F] * 6 + [_D, _D, _I3, _I3, _I3, _V],
    "tb200_smagorinsky_step": [_F] * 7 + [_D, _D, _D, _D, _I3, _I3, _I3, _V],
    "tb200_implicit_vertical_advection": [_I] + [_F] * 13 + [_D, _D, _I3, _I3, _V],
    "tb200_vertical_advection": [_I, _I] + [_F] * 13 + [_D, C.c_uint32, _I3, _I3, _V],
    "tb200_vertical_advection_step": [_I, _I, _F, _I, _FPP, _FPP, _FPP, _D, _D, _I3, _I3, _V],
    "tb200_halo_pack": [_FPP, _I, C.c_void_p, _I3, _I3, _V],
    "tb200_halo_unpack": [_FPP, _I, C.c_void_p, _I3, _I3, _V],
    "tb200_halo_push": [_FPP, _I, C.c_void_p, _I, _V, _I, _I, _V],
    "tb200_halo_pull": [_FPP, _I, C.c_void_p, _I, _V, _I, _I, _V],
    "tb200_p2p_alloc": [C.c_size_t, C.POINTER(C.c_void_p)],
    "tb200_p2p_free": [_V],
    "tb200_p2p_export": [_V, C.c_char_p],
    "tb200_p2p_import": [C.c_char_p, C.POINTER(C.c_void_p)],
    "tb200_p2p_release": [_V],
    "tb200_p2p_channel_error": [_V, C.POINTER(C.c_int)],
    "tb200_selftest_division": [C.c_uint64, C.c_uint64, _V, _V],
    "tb200_ctx_create": [C.POINTER(C.c_void_p)],
    "tb200_ctx_destroy": [_V],
    "tb200_ctx_scratch": [_V, C.POINTER(C.c_int64), _I, FieldP],
}

KESSLER_FLAGS = {"p_on_interfaces": 1, "rain_evaporation": 2, "ow_qc": 4, "ow_qr": 8, "ow_qv": 16,
                 "ow_theta": 32}

_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load the shared library (building it first if it is absent and nvcc is around)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and build_if_missing:
        from tasmania_b200 import build as _build

        _build.build()
    if not os.path.exists(LIB_PATH):
        raise B200Error(
            f"{LIB_PATH} not found: the b200 backend has no CPU fallback; build it with "
            "`python -m tasmania_b200.build`"
        )
    lib = C.CDLL(LIB_PATH)
    lib.tb200_last_error.restype = C.c_char_p
    lib.tb200_last_error.argtypes = []
    lib.tb200_version.restype = C.c_int
    lib.tb200_device_count.restype = C.c_int
    lib.tb200_launch_count.restype = C.c_longlong
    lib.tb200_launch_count.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


def exported_symbols():
    return ["tb200_last_error", "tb200_version", "tb200_device_count", "tb200_launch_count",
            *SIGNATURES]


def launch_count() -> int:
    """Kernels launched by the library so far (``tb200_launch_count``)."""
    return int(load().tb200_launch_count())


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().tb200_last_error().decode("utf-8", "replace")
        raise B200Error(f"{what} failed (code {rc}): {msg}")


def current_stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def as_field(x) -> Optional[Field]:
    """Build a ``tb200_field`` from anything exposing ``__cuda_array_interface__`` (B200Array,
    torch CUDA tensor, cupy array).  3-D fp64 only; lower ranks are right-padded with 1s."""
    if x is None:
        return None
    cai = getattr(x, "__cuda_array_interface__", None)
    if cai is None:
        raise B200Error(
            f"{type(x).__name__} is not a device array: the b200 backend computes on GPU memory "
            "only (no CPU fallback); allocate with tasmania_b200.zeros / as_storage"
        )
    if cai["typestr"] not in ("<f8", "=f8", "|f8"):
        raise B200Error(f"b200 kernels are fp64; got typestr {cai['typestr']}")
    shape = tuple(cai["shape"])
    if len(shape) > 3:
        raise B200Error(f"fields are at most 3-D, got shape {shape}")
    strides = cai.get("strides")
    if strides is None:
        strides, acc = [], 8
        for n in reversed(shape):
            strides.insert(0, acc)
            acc *= n
    strides = [s // 8 for s in strides]
    while len(shape) < 3:
        shape = shape + (1,)
        strides = strides + [0]
    f = Field()
    f.ptr = cai["data"][0]
    f.shape[:] = shape
    f.stride[:] = strides
    return f


class ScratchField:
    """A library-owned scratch field (``tb200_ctx_scratch``) as a device array the marshalling code
    accepts: ``__cuda_array_interface__`` only -- scratch is never read on the host."""

    def __init__(self, field: Field, owner):
        self.shape = tuple(int(n) for n in field.shape)
        self._strides = tuple(int(s) * 8 for s in field.stride)
        self._ptr = int(field.ptr)
        self._owner = owner  # keeps the context alive

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": "<f8", "data": (self._ptr, False), "version": 3,
                "strides": self._strides}


class Context:
    """``tb200_ctx``: owner of the fused kernels' scratch memory, held by the host-side object that
    issues the fused calls; freed with the object."""

    def __init__(self):
        h = C.c_void_p()
        check(load().tb200_ctx_create(C.byref(h)), "tb200_ctx_create")
        self._h = h

    def scratch(self, shape, count):
        fields = (Field * count)()
        check(load().tb200_ctx_scratch(self._h, (C.c_int64 * 3)(*[int(n) for n in shape]), count, fields),
              "tb200_ctx_scratch")
        return tuple(ScratchField(fields[n], self) for n in range(count))

    def __copy__(self):
        raise TypeError("a tb200_ctx owns device memory: not copyable")

    def __deepcopy__(self, memo):
        raise TypeError("a tb200_ctx owns device memory: not copyable")

    def __reduce__(self):
        raise TypeError("a tb200_ctx owns device memory: not picklable")

    def __del__(self):
        try:
            if self._h:
                load().tb200_ctx_destroy(self._h)
                self._h = None
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


def fp(x) -> Optional[FieldP]:
    f = as_field(x)
    return None if f is None else C.pointer(f)


def int3(t: Sequence[int]):
    return Int3(int(t[0]), int(t[1]), int(t[2]))
