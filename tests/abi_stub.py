# -*- coding: utf-8 -*-
"""Host-logic test double for libtasmania_b200.so (test infrastructure, CPU only).

The product has no CPU path: its kernels live behind the C ABI and refuse host memory.  To test
the *host* side (couplers, model assembly, argument marshalling) in the GPU-less container, this
module swaps the loaded library for a stub whose entry points

  * have the ctypes prototypes of ``tasmania_b200.lib.SIGNATURES`` -- every call is type-checked
    against the declared C signature exactly as a real call would be -- and are recorded;
  * do nothing, except ``tb200_fma_fields``, ``tb200_elementwise`` (copy / fma) and
    ``tb200_relax_frame``, which are carried out with numpy on the host buffers so that the
    arithmetic of the coupling layer (stage factors, buffer swaps, accumulate-or-overwrite) and of
    the boundary glue can be followed end to end.

Storages are allocated on the host for the duration (``storage.DEFAULT_DEVICE_OVERRIDE``).
Nothing in tasmania_b200 knows about this file.
"""
import contextlib
import ctypes as C

import numpy as np

from tasmania_b200 import lib, stencils, storage


def _host_field(x):
    if x is None:
        return None
    t = x.t if hasattr(x, "t") else x
    shape, strides = tuple(t.shape), list(t.stride())
    while len(shape) < 3:
        shape, strides = shape + (1,), strides + [0]
    f = lib.Field()
    f.ptr = t.data_ptr()
    f.shape[:] = shape
    f.stride[:] = strides
    return f


def _as_numpy(field):
    shape = tuple(int(n) for n in field.shape)
    stride = tuple(int(s) for s in field.stride)
    n = 1 + sum((a - 1) * s for a, s in zip(shape, stride))
    flat = np.ctypeslib.as_array((C.c_double * n).from_address(field.ptr))
    return np.lib.stride_tricks.as_strided(flat, shape, tuple(8 * s for s in stride))


def _box(o, d):
    return tuple(slice(int(o[n]), int(o[n]) + int(d[n])) for n in range(3))


def _freeze(name, argtypes, args):
    """Copy the ctypes arguments of a call so that it can be re-issued later (the marshalling
    code frees its tb200_field structs right after the call) and describe it for traces."""
    frozen, desc, keep = [], [], []
    nf = {"tb200_fma_fields": lambda a: a[0], "tb200_relax_frame": lambda a: a[0],
          "tb200_vertical_advection_step": lambda a: a[3], "tb200_halo_pack": lambda a: a[1],
          "tb200_halo_unpack": lambda a: a[1]}.get(name, lambda a: 3)(args)
    for t, v in zip(argtypes, args):
        if t is lib.FieldP:
            if not v:
                frozen.append(None), desc.append(None)
            else:
                f = lib.Field.from_buffer_copy(v.contents)
                keep.append(f)
                frozen.append(C.pointer(f))
                desc.append((f.ptr, tuple(f.shape), tuple(f.stride)))
        elif t is C.POINTER(lib.FieldP):
            if not v:
                frozen.append(None), desc.append(None)
            else:
                fs = [lib.Field.from_buffer_copy(v[m].contents) if v[m] else None for m in range(nf)]
                keep.append(fs)
                arr = (lib.FieldP * nf)(*[C.pointer(f) if f is not None else None for f in fs])
                keep.append(arr)
                frozen.append(arr)
                desc.append(tuple(None if f is None else (f.ptr, tuple(f.shape), tuple(f.stride))
                                  for f in fs))
        elif t is C.POINTER(C.c_int32):
            n = 3  # origin / domain triplets, except the two arrays of tb200_relax_frame
            if name == "tb200_relax_frame":
                n = 3 * args[0] if len(frozen) == 4 else 4
            vals = (C.c_int32 * n)(*[v[m] for m in range(n)])
            keep.append(vals)
            frozen.append(vals), desc.append(tuple(vals))
        elif t is C.POINTER(lib.StageCfg):  # configuration block of the fused stage
            cfg = lib.StageCfg.from_buffer_copy(v.contents)
            keep.append(cfg)
            frozen.append(C.pointer(cfg))
            row = []
            for f, ftype in lib.StageCfg._fields_:
                val = getattr(cfg, f)
                if ftype is lib.FieldP:  # optional fields of the block (slow tendencies): copy the struct
                    if val:
                        fld = lib.Field.from_buffer_copy(val.contents)
                        keep.append(fld)
                        setattr(cfg, f, C.pointer(fld))
                        row.append((fld.ptr, tuple(fld.shape), tuple(fld.stride)))
                    else:
                        row.append(None)
                else:
                    row.append(tuple(val) if hasattr(val, "__len__") else val)
            desc.append(tuple(row))
        elif t is C.POINTER(C.c_double):  # the four physical constants of the column scans
            vals = (C.c_double * 4)(*[v[m] for m in range(4)]) if v else None
            keep.append(vals)
            frozen.append(vals), desc.append(tuple(vals) if vals is not None else None)
        elif t in (C.c_double, C.c_int, C.c_uint32):
            frozen.append(v), desc.append(v)
        else:  # stream handles, config structs, raw buffers: passed through, not traced
            frozen.append(v), desc.append("*")
    return frozen, tuple(desc), keep


class AbiStub:
    def __init__(self):
        self.calls = []
        self.trace = None      # list of (name, argument description) of every EXECUTED call
        self.recording = None  # list of frozen calls while a fake graph capture is on
        self._cb = {}
        for name, argtypes in lib.SIGNATURES.items():
            self._cb[name] = C.CFUNCTYPE(C.c_int, *argtypes)(self._make(name, argtypes))

    def _make(self, name, argtypes):
        if name == "tb200_stage_lazy_velocities":  # a query, not a launch: the default kernel path
            return lambda nz: 1 if nz <= 64 else 0

        def fn(*args):
            self.calls.append(name)
            if self.recording is not None:  # stream capture: record, do not execute
                self.recording.append((name,) + _freeze(name, argtypes, args))
                return 0
            self._execute(name, args, _freeze(name, argtypes, args)[1] if self.trace is not None else None)
            return 0

        return fn

    def _execute(self, name, args, desc):
        if self.trace is not None:
            self.trace.append((name, desc))
        handler = getattr(self, "_do_" + name, None)
        if handler is not None:
            handler(*args)

    def __getattr__(self, name):
        try:
            return self.__dict__["_cb"][name]
        except KeyError:
            raise AttributeError(name) from None

    def tb200_last_error(self):
        return b"stub"

    def tb200_launch_count(self):
        return len(self.calls)

    # ---- the two entry points the stub really executes
    def _do_tb200_fma_fields(self, n, outs, a, b, f, o, d, stream):
        box = _box(o, d)
        for m in range(n):
            _as_numpy(outs[m].contents)[box] = (
                _as_numpy(a[m].contents)[box] + f * _as_numpy(b[m].contents)[box])

    def _do_tb200_elementwise(self, op, out, a, b, c, f, o, d, stream):
        box = _box(o, d)
        if op == lib.ELEMENTWISE_OPS["copy"]:
            _as_numpy(out.contents)[box] = _as_numpy(a.contents)[box]
        elif op == lib.ELEMENTWISE_OPS["fma"]:
            _as_numpy(out.contents)[box] = _as_numpy(a.contents)[box] + f * _as_numpy(b.contents)[box]

    def _do_tb200_relax_frame(self, n, phi, ref, gamma, extents, interior, stream):
        """Relaxed.enforce_raw on the frame outside ``interior`` (algorithms.py:L32-L43 per point)."""
        g2d = _as_numpy(gamma.contents)[:, :, 0]
        i_lo, i_hi, j_lo, j_hi = (int(interior[m]) for m in range(4))
        for m in range(n):
            mi, mj, mk = (int(extents[3 * m + c]) for c in range(3))
            frame = np.ones((mi, mj), dtype=bool)
            frame[i_lo:i_hi, j_lo:j_hi] = False
            g = g2d[:mi, :mj, None]
            a = _as_numpy(phi[m].contents)[:mi, :mj, :mk]
            r = _as_numpy(ref[m].contents)[:mi, :mj, :mk]
            relaxed = np.where(g == 1.0, r, a - g * (a - r))
            a[...] = np.where(frame[:, :, None] & (g != 0.0), relaxed, a)

    def count(self, name):
        return sum(1 for c in self.calls if c == name)


class FakeCapture:
    """Stands in for torch.cuda.CUDAGraph in tasmania_b200.graphs.GraphedLoop: ``capture`` records
    the ABI calls ``fn`` issues WITHOUT executing them (like stream capture), ``replay`` re-issues
    them with the recorded arguments.  Slice assignments on storages (torch copy kernels on the
    device, which stream capture records as well) are deferred in the same way."""

    stub = None  # set by the test

    def capture(self, fn):
        stub = self.stub
        assert stub.recording is None
        stub.recording = []
        setitem = storage.B200Array.__setitem__

        def deferred_setitem(array, idx, value):
            stub.recording.append(("<setitem>", lambda: setitem(array, idx, value), None, None))

        storage.B200Array.__setitem__ = deferred_setitem
        try:
            fn()
        finally:
            storage.B200Array.__setitem__ = setitem
            self.recorded, stub.recording = stub.recording, None

    def replay(self):
        for name, frozen, desc, _keep in self.recorded:
            if name == "<setitem>":
                frozen()
            else:
                self.stub._execute(name, frozen, desc)


def canonical(trace):
    """A trace with every pointer replaced by the order of its first appearance, so that runs of two
    model instances (different allocations, same allocation order) can be compared."""
    ids = {}

    def canon(x):
        if isinstance(x, tuple) and len(x) == 3 and isinstance(x[1], tuple) and isinstance(x[0], int):
            return (ids.setdefault(x[0], len(ids)),) + x[1:]
        if isinstance(x, tuple):
            return tuple(canon(y) for y in x)
        return x

    return [(name, canon(desc)) for name, desc in trace]


@contextlib.contextmanager
def stubbed_library(stub_cls=None):
    """``with stubbed_library() as stub:`` -- host storages + the recording ABI stub (or a subclass
    of it, e.g. tests/abi_oracle.py:OracleStub, which carries every call out with the oracle)."""
    saved = (lib._lib, lib.as_field, lib.current_stream, stencils._f, storage.DEFAULT_DEVICE_OVERRIDE)
    stub = (stub_cls or AbiStub)()
    lib._lib, lib.as_field, lib.current_stream = stub, _host_field, lambda: 0
    stencils._f = _host_field
    storage.DEFAULT_DEVICE_OVERRIDE = "cpu"
    try:
        yield stub
    finally:
        (lib._lib, lib.as_field, lib.current_stream, stencils._f,
         storage.DEFAULT_DEVICE_OVERRIDE) = saved
