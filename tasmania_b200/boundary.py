# -*- coding: utf-8 -*-
"""Lateral boundary conditions on b200 storages (row K5 of SURVEY.md section 8a): the raw-field
operations of the reference's ``HorizontalBoundary`` subclasses

  Relaxed    src/tasmania/domain/subclasses/horizontal_boundaries/relaxed.py:L34-L247
  Periodic   src/tasmania/domain/subclasses/horizontal_boundaries/periodic.py:L32-L122
  Dirichlet  src/tasmania/domain/subclasses/horizontal_boundaries/dirichlet.py:L40-L160
  Identity   src/tasmania/domain/subclasses/horizontal_boundaries/identity.py:L30-L79
  Relaxed1DX / 1DY   relaxed.py:L250-L461, L463-L678  (grids with ny == 1 / nx == 1)
  Periodic1DX / 1DY  periodic.py:L125-L214, L217-L306
  enforce_raw  src/tasmania/domain/horizontal_boundary.py:L299-L344

with the same method names.  Coefficients are built once on the host (numpy, O(nx ny)) and
uploaded; every field operation is a CUDA kernel.
"""
from __future__ import annotations

import os

import numpy as np

from tasmania_b200 import lib, storage
from tasmania_b200.framework import BackendOptions, StencilFactory, StorageOptions


def _extent(nx, ny, nz, name):
    name = name or ""
    mi = nx + 1 if ("at_u_locations" in name or "at_uv_locations" in name) else nx
    mj = ny + 1 if ("at_v_locations" in name or "at_uv_locations" in name) else ny
    mk = nz + 1 if "on_interface_levels" in name else nz
    return mi, mj, mk


class HorizontalBoundary(StencilFactory):
    """Common part: sizes, the reference state (deep copies), ``enforce_raw``."""

    type = None

    def __init__(self, nx, ny, nz, nb, backend_options=None, storage_options=None):
        super().__init__("b200", backend_options or BackendOptions(), storage_options or StorageOptions())
        self.nx, self.ny, self.nz, self.nb = int(nx), int(ny), int(nz), int(nb)
        self._ref_state = None

    @property
    def reference_state(self):
        return self._ref_state if self._ref_state is not None else {}

    @reference_state.setter
    def reference_state(self, ref_state):
        # deep copies, horizontal_boundary.py:L131-L144
        self._ref_state = {}
        for name, val in ref_state.items():
            if name == "time":
                self._ref_state[name] = val
            else:
                self._ref_state[name] = storage.as_storage(
                    val, device=self.storage_options.device).copy()

    @staticmethod
    def factory(boundary_type, nx, ny, nz, nb, **kwargs):
        classes = {"relaxed": Relaxed, "periodic": Periodic, "dirichlet": Dirichlet,
                   "identity": Identity}
        if boundary_type not in classes:
            raise ValueError(f"unknown (or out-of-scope) horizontal boundary type {boundary_type!r}")
        # the reference's dispatch on degenerate grids (relaxed.py:L680-L710, periodic.py:L309-L322)
        one_d = {"relaxed": (Relaxed1DX, Relaxed1DY), "periodic": (Periodic1DX, Periodic1DY)}
        if boundary_type in one_d and (nx == 1 or ny == 1):
            return one_d[boundary_type][0 if ny == 1 else 1](nx, ny, nz, nb, **kwargs)
        return classes[boundary_type](nx, ny, nz, nb, **kwargs)

    def enforce_raw(self, state, field_properties=None):
        """Only fields that have a reference value are touched (base-class behaviour)."""
        ref = self.reference_state
        for name in state:
            if name == "time" or name not in ref:
                continue
            if field_properties is not None and name not in field_properties:
                continue
            self.enforce_field(state[name], field_name=name, time=state.get("time"))

    def enforce_field(self, field, field_name=None, field_units=None, time=None):
        raise NotImplementedError


def relaxed_gamma_window(global_extent, offset, shape, rel):
    """The relaxation coefficients of relaxed.py:L193-L247 on a window of the global grid, in
    closed form: a point of ring ``m = min(i, j, NX-1-i, NY-1-j)`` gets ``rel[m]`` (corners take
    the coefficient of the outer ring), the interior 0, and the extra staggered row / column
    ``i == NX`` / ``j == NY`` gets 1.  Identical to the block-wise construction (checked in
    tests/test_distributed_cpu.py)."""
    nxg, nyg = global_extent
    nr = len(rel)
    i = np.arange(shape[0])[:, None] + offset[0]
    j = np.arange(shape[1])[None, :] + offset[1]
    ring = np.minimum(np.minimum(i, nxg - 1 - i), np.minimum(j, nyg - 1 - j))
    inside = (i < nxg) & (j < nyg)
    relx = np.concatenate((np.asarray(rel, dtype=float), [0.0]))
    g = relx[np.clip(ring, 0, nr)]
    g = np.where(inside, g, 0.0)
    g = np.where(((i == nxg) & (j <= nyg)) | ((j == nyg) & (i <= nxg)), 1.0, g)
    return g


def relax_frame(fields, refs, extents, gamma, free_box):
    """``tb200_relax_frame``: relax all ``fields`` towards ``refs`` on the frame outside
    ``free_box`` (a box on which ``gamma`` vanishes), eight fields per launch."""
    import ctypes as C

    box = (C.c_int32 * 4)(*free_box)
    g = lib.as_field(gamma)
    for lo in range(0, len(fields), 8):
        phi = [lib.as_field(f) for f in fields[lo:lo + 8]]
        ref = [lib.as_field(f) for f in refs[lo:lo + 8]]
        ext = (C.c_int32 * (3 * len(phi)))(*[int(v) for e in extents[lo:lo + 8] for v in e])
        lib.check(lib.load().tb200_relax_frame(
            len(phi), (lib.FieldP * len(phi))(*[C.pointer(f) for f in phi]),
            (lib.FieldP * len(phi))(*[C.pointer(f) for f in ref]), g, ext, box,
            lib.current_stream()), "tb200_relax_frame")


class Relaxed(HorizontalBoundary):
    """Relaxed boundary conditions (``ni = nx``, ``nj = ny``)."""

    type = "relaxed"

    def __init__(self, nx, ny, nz, nb, nr=8, backend_options=None, storage_options=None,
                 storage_shape=None, global_extent=None, offset=(0, 0)):
        """``global_extent=(NX, NY)`` and ``offset=(i0, j0)`` describe a sub-domain of a 2-D
        decomposed grid (tasmania_b200.distributed): the object then handles the local
        ``nx x ny`` window whose point (0, 0) is the global point (i0, j0), with the relaxation
        coefficients of the *global* boundary."""
        assert nx > 1 and ny > 1
        self._global = tuple(global_extent) if global_extent is not None else (int(nx), int(ny))
        self._offset = (int(offset[0]), int(offset[1]))
        nxg, nyg = self._global
        assert nr <= nxg / 2 and nr <= nyg / 2, "Depth of relaxation region cannot exceed n/2."
        assert nr <= 8, "Depth of relaxation region cannot exceed 8."
        assert nb <= nr, "Number of boundary layers cannot exceed depth of relaxation region."
        super().__init__(nx, ny, nz, nb, backend_options, storage_options)
        self.nr = nr
        self.ni, self.nj = self.nx, self.ny
        self._shape = tuple(storage_shape or (nx + 1, ny + 1, nz + 1))
        self._allocate_coefficient_matrix()
        self._stencil = self.compile_stencil("irelax")

    def _allocate_coefficient_matrix(self):
        """relaxed.py:L193-L247; gamma does not depend on k -> stored once as (ni, nj, 1)
        and broadcast through a zero k-stride view (saves a 3-D storage the reference streams
        on every call)."""
        nx, ny, nb, nr = self.nx, self.ny, self.nb, self.nr
        rel = np.array([1.0] + [1.0 - np.tanh(0.5 * m) for m in range(1, 8)])[:nr]
        rel[:nb] = 1.0
        if self._global != (nx, ny) or self._offset != (0, 0):
            g = relaxed_gamma_window(self._global, self._offset, self._shape[:2], rel)
            g2d = storage.as_storage(g[:, :, None], device=self.storage_options.device)
            self._gamma2d = g2d
            self._gamma = storage.B200Array(g2d.t.expand(-1, -1, self._shape[2]))
            self._free_box = self._gamma_free_box(g)
            return
        rrel = rel[::-1]
        g = np.zeros((self._shape[0], self._shape[1]))
        corner = np.zeros((nr, nr))
        for i in range(nr):
            corner[i, i:] = rel[i]
            corner[i:, i] = rel[i]
        g[:nr, :nr] = corner
        g[:nr, nr : ny - nr] = rel[:, None]
        g[:nr, ny - nr : ny] = corner[:, ::-1]
        g[nx - nr : nx, :nr] = corner[::-1, :]
        g[nx - nr : nx, nr : ny - nr] = rrel[:, None]
        g[nx - nr : nx, ny - nr : ny] = corner[::-1, ::-1]
        g[nr : nx - nr, :nr] = rel[None, :]
        g[nr : nx - nr, ny - nr : ny] = rrel[None, :]
        g[nx : nx + 1, : ny + 1] = 1.0
        g[: nx + 1, ny : ny + 1] = 1.0
        g2d = storage.as_storage(g[:, :, None], device=self.storage_options.device)
        self._gamma2d = g2d
        self._free_box = self._gamma_free_box(g)
        # (ni, nj, nk) view with stride 0 along k
        self._gamma = storage.B200Array(g2d.t.expand(-1, -1, self._shape[2]))

    def enforce_field(self, field, field_name=None, field_units=None, time=None):
        mi, mj, mk = _extent(self.nx, self.ny, self.nz, field_name)
        self._stencil(in_gamma=self._gamma, in_phi_ref=self.reference_state[field_name],
                      inout_phi=field, origin=(0, 0, 0), domain=(mi, mj, mk))

    @staticmethod
    def _gamma_free_box(g):
        """(i_lo, i_hi, j_lo, j_hi): a box of the (nx + 1, ny + 1) coefficient matrix on which gamma
        vanishes -- everything but the nr outer rings and the staggered extra row / column on a
        whole domain, everything on an interior sub-domain of a decomposed grid.  Found from the
        matrix itself and verified; (0, 0, 0, 0) (= relax everywhere) if the zeros are not a box."""
        nz_mask = np.asarray(g) != 0.0
        rows = np.flatnonzero(~nz_mask.all(axis=1))  # rows with at least one zero
        cols = np.flatnonzero(~nz_mask.all(axis=0))
        if rows.size and cols.size:
            # the zero set of a ring-structured matrix is the box spanned by the zeros of its
            # middle row and column
            im, jm = int(rows[rows.size // 2]), int(cols[cols.size // 2])
            zi, zj = np.flatnonzero(~nz_mask[:, jm]), np.flatnonzero(~nz_mask[im, :])
            if zi.size and zj.size:
                cand = (int(zi[0]), int(zi[-1]) + 1, int(zj[0]), int(zj[-1]) + 1)
                if not nz_mask[cand[0]:cand[1], cand[2]:cand[3]].any():
                    return cand
        return (0, 0, 0, 0)

    def enforce_raw(self, state, field_properties=None):
        """All fields with a reference value in one launch over the frame where gamma != 0
        (``tb200_relax_frame``) instead of one full-box ``irelax`` per field; TB200_RELAX=full
        keeps the per-field path (bit-identical, tests/test_gpu_stencils.py)."""
        if os.environ.get("TB200_RELAX", "frame") != "frame":
            return super().enforce_raw(state, field_properties)
        ref = self.reference_state
        names = [n for n in state if n != "time" and n in ref
                 and (field_properties is None or n in field_properties)]
        if not names:
            return
        relax_frame([state[n] for n in names], [ref[n] for n in names],
                    [_extent(self.nx, self.ny, self.nz, n) for n in names], self._gamma2d, self._free_box)

    def _outermost(self, axis, field, field_name):
        mi, mj, _ = _extent(self.nx, self.ny, self.nz, field_name)
        lib.check(lib.load().tb200_set_outermost_layers(
            lib.as_field(field), lib.as_field(self.reference_state[field_name]), axis, mi, mj,
            lib.current_stream()), "tb200_set_outermost_layers")

    def set_outermost_layers_x(self, field, field_name=None, field_units=None, time=None):
        self._outermost(0, field, field_name)

    def set_outermost_layers_y(self, field, field_name=None, field_units=None, time=None):
        self._outermost(1, field, field_name)

    def get_numerical_field(self, field, field_name=None):
        return field

    def get_physical_field(self, field, field_name=None):
        return field


class Periodic(HorizontalBoundary):
    """Periodic conditions; fields live on the numerical grid ``(nx + 2 nb, ny + 2 nb)``."""

    type = "periodic"

    def __init__(self, nx, ny, nz, nb, backend_options=None, storage_options=None):
        assert nx > 1 and ny > 1 and nb <= nx / 2 and nb <= ny / 2
        super().__init__(nx, ny, nz, nb, backend_options, storage_options)
        self.ni, self.nj = self.nx + 2 * nb, self.ny + 2 * nb

    def enforce_field(self, field, field_name=None, field_units=None, time=None):
        mx, my, _ = _extent(self.nx, self.ny, self.nz, field_name)
        lib.check(lib.load().tb200_periodic_enforce(
            lib.as_field(field), self.nx, self.ny, self.nb, mx, my, lib.current_stream()),
            "tb200_periodic_enforce")

    def get_numerical_field(self, field, field_name=None):
        """periodic.py:L64-L96"""
        nb = self.nb
        mx, my, _ = _extent(self.nx, self.ny, self.nz, field_name)
        src = storage.as_storage(field, device=self.storage_options.device)
        shape = (src.shape[0] + 2 * nb, src.shape[1] + 2 * nb) + tuple(src.shape[2:])
        trg = self.zeros(shape=shape)
        trg[nb : mx + nb, nb : my + nb] = src[:mx, :my]
        self.enforce_field(trg, field_name)
        return trg

    def get_physical_field(self, field, field_name=None):
        return field[self.nb : -self.nb, self.nb : -self.nb]

    def set_outermost_layers_x(self, field, field_name=None, field_units=None, time=None):
        field[0, :] = field[-2, :]
        field[-1, :] = field[1, :]

    def set_outermost_layers_y(self, field, field_name=None, field_units=None, time=None):
        field[:, 0] = field[:, -2]
        field[:, -1] = field[:, 1]


class Dirichlet(HorizontalBoundary):
    """Dirichlet conditions: ``core(time, grid, slice_x, slice_y, field_name, field_units)``
    provides the rim values.  As in the reference the core is host code (numpy) and the four
    rim slabs are uploaded on every call (dirichlet.py:L98-L150)."""

    type = "dirichlet"

    def __init__(self, nx, ny, nz, nb, core=None, grid=None, backend_options=None,
                 storage_options=None):
        assert nx > 1 and ny > 1 and nb <= nx / 2 and nb <= ny / 2
        super().__init__(nx, ny, nz, nb, backend_options, storage_options)
        self.ni, self.nj = self.nx, self.ny
        self.core, self.grid = core, grid

    def enforce_field(self, field, field_name=None, field_units=None, time=None):
        nb, core, g = self.nb, self.core, self.grid
        mi, mj, _ = _extent(self.nx, self.ny, self.nz, field_name)
        for sx, sy in (
            (slice(0, nb), slice(0, mj)),
            (slice(mi - nb, mi), slice(0, mj)),
            (slice(nb, mi - nb), slice(0, nb)),
            (slice(nb, mi - nb), slice(mj - nb, mj)),
        ):
            vals = np.asarray(core(time, g, sx, sy, field_name, field_units))
            field[sx, sy, : vals.shape[2]] = vals

    def set_outermost_layers_x(self, field, field_name=None, field_units=None, time=None):
        """dirichlet.py:L161-L190: the two outermost x-layers from the core (host code, uploaded)."""
        core, g = self.core, self.grid
        mi, mj, _ = _extent(self.nx, self.ny, self.nz, field_name)
        for sx in (slice(0, 1), slice(mi - 1, mi)):
            vals = np.asarray(core(time, g, sx, slice(0, mj), field_name, field_units))
            field[sx, slice(0, mj), : vals.shape[2]] = vals

    def set_outermost_layers_y(self, field, field_name=None, field_units=None, time=None):
        """dirichlet.py:L192-L220."""
        core, g = self.core, self.grid
        mi, mj, _ = _extent(self.nx, self.ny, self.nz, field_name)
        for sy in (slice(0, 1), slice(mj - 1, mj)):
            vals = np.asarray(core(time, g, slice(0, mi), sy, field_name, field_units))
            field[slice(0, mi), sy, : vals.shape[2]] = vals

    def get_numerical_field(self, field, field_name=None):
        return field

    def get_physical_field(self, field, field_name=None):
        return field


class Identity(HorizontalBoundary):
    """No-op conditions: the numerical grid is the physical grid and no field is touched
    (identity.py:L30-L79) -- no kernel, no launch."""

    type = "identity"

    def __init__(self, nx, ny, nz, nb, backend_options=None, storage_options=None):
        assert nx > 1 and ny > 1 and nb <= nx / 2 and nb <= ny / 2
        super().__init__(nx, ny, nz, nb, backend_options, storage_options)
        self.ni, self.nj = self.nx, self.ny

    def get_numerical_field(self, field, field_name=None):
        return field

    def get_physical_field(self, field, field_name=None):
        return field

    def enforce_field(self, field, field_name=None, field_units=None, time=None):
        pass

    def set_outermost_layers_x(self, field, field_name=None, field_units=None, time=None):
        pass

    def set_outermost_layers_y(self, field, field_name=None, field_units=None, time=None):
        pass


# ---------------------------------------------------------------- one-dimensional grids
class _OneDimensional(HorizontalBoundary):
    """Shared by the ...1DX (ny == 1) and ...1DY (nx == 1) variants: ``axis`` is the axis the
    grid extends along; the degenerate axis carries 2 nb + 1 copies of its single point on the
    numerical grid.  The slab copies are slice assignments on b200 storages (device-to-device
    copies), exactly what the reference's own classes do on this backend through the plugin."""

    axis = 0

    def _n(self):
        return self.nx if self.axis == 0 else self.ny

    def _ix(self, along, across, k=None):
        """index tuple with ``along`` on the grid's axis and ``across`` on the degenerate one"""
        ij = (along, across) if self.axis == 0 else (across, along)
        return ij if k is None else ij + (k,)

    def _staggered(self, name):
        """(staggered along the axis, staggered across it)"""
        name = name or ""
        u = "at_u_locations" in name or "at_uv_locations" in name
        v = "at_v_locations" in name or "at_uv_locations" in name
        return (u, v) if self.axis == 0 else (v, u)


class _Relaxed1D(_OneDimensional):
    type = "relaxed"

    def __init__(self, nx, ny, nz, nb, nr=8, backend_options=None, storage_options=None,
                 storage_shape=None):
        super().__init__(nx, ny, nz, nb, backend_options, storage_options)
        n = self._n()
        assert n > 1 and (ny if self.axis == 0 else nx) == 1
        assert nr <= n / 2 and nr <= 8 and nb <= nr
        self.nr = nr
        self.ni, self.nj = (self.nx, 2 * nb + 1) if self.axis == 0 else (2 * nb + 1, self.ny)
        self._shape = tuple(storage_shape or (self.ni + 1, self.nj + 1, nz + 1))
        # relaxed.py:L424-L461 / L641-L678: the coefficients on the two middle lines only
        rel = np.array([1.0] + [1.0 - np.tanh(0.5 * m) for m in range(1, 8)])[:nr]
        rel[:nb] = 1.0
        g = np.zeros(self._shape[:2])
        mid = slice(nb, nb + 2)
        g[self._ix(slice(0, nr), mid)] = rel[:, None] if self.axis == 0 else rel[None, :]
        g[self._ix(slice(n - nr, n), mid)] = rel[::-1][:, None] if self.axis == 0 else rel[::-1][None, :]
        g[self._ix(slice(n, n + 1), mid)] = 1.0
        self._gamma2d = storage.as_storage(g[:, :, None], device=self.storage_options.device)
        self._gamma = storage.B200Array(self._gamma2d.t.expand(-1, -1, self._shape[2]))
        self._stencil = self.compile_stencil("irelax")

    def _extents(self, name):
        return _extent(self.ni, self.nj, self.nz, name)

    def get_numerical_field(self, field, field_name=None):
        nb = self.nb
        src = storage.as_storage(field, device=self.storage_options.device)
        shape = list(src.shape)
        shape[1 - self.axis] += 2 * nb
        trg = self.zeros(shape=tuple(shape))
        trg[self._ix(slice(None), slice(None, nb + 1))] = src[self._ix(slice(None), slice(0, 1))]
        trg[self._ix(slice(None), slice(nb + 1, None))] = src[self._ix(slice(None), slice(-1, None))]
        return trg

    def enforce_field(self, field, field_name=None, field_units=None, time=None):
        nb = self.nb
        m = self._extents(field_name)
        ma, mo, mk = m[self.axis], m[1 - self.axis], m[2]
        origin, domain = [0, 0, 0], [m[0], m[1], mk]
        origin[1 - self.axis], domain[1 - self.axis] = nb, mo - nb
        self._stencil(in_gamma=self._gamma, in_phi_ref=self.reference_state[field_name],
                      inout_phi=field, origin=tuple(origin), domain=tuple(domain))
        # repeat the innermost line(s) across the degenerate direction
        k, al = slice(0, mk), slice(0, ma)
        field[self._ix(al, slice(0, nb), k)] = field[self._ix(al, slice(nb, nb + 1), k)]
        field[self._ix(al, slice(mo - nb, mo), k)] = field[self._ix(al, slice(mo - nb - 1, mo - nb), k)]

    def get_physical_field(self, field, field_name=None):
        return field[self._ix(slice(None), slice(self.nb, -self.nb))]

    def _outermost(self, x_layers, field, field_name):
        mi, mj, _ = self._extents(field_name)
        ref = self.reference_state[field_name]
        if x_layers:
            field[0, :mj] = ref[0, :mj]
            field[mi - 1, :mj] = ref[mi - 1, :mj]
        else:
            field[:mi, 0] = ref[:mi, 0]
            field[:mi, mj - 1] = ref[:mi, mj - 1]

    def set_outermost_layers_x(self, field, field_name=None, field_units=None, time=None):
        self._outermost(True, field, field_name)

    def set_outermost_layers_y(self, field, field_name=None, field_units=None, time=None):
        self._outermost(False, field, field_name)


class Relaxed1DX(_Relaxed1D):
    """relaxed.py:L250-L461 (ny == 1)."""

    axis = 0


class Relaxed1DY(_Relaxed1D):
    """relaxed.py:L463-L678 (nx == 1)."""

    axis = 1


class _Periodic1D(_OneDimensional):
    type = "periodic"

    def __init__(self, nx, ny, nz, nb, backend_options=None, storage_options=None):
        super().__init__(nx, ny, nz, nb, backend_options, storage_options)
        n = self._n()
        assert n > 1 and (ny if self.axis == 0 else nx) == 1 and nb <= n / 2
        self.ni, self.nj = (self.nx + 2 * nb, 2 * nb + 1) if self.axis == 0 else (2 * nb + 1, self.ny + 2 * nb)

    def enforce_field(self, field, field_name=None, field_units=None, time=None):
        """periodic.py:L190-L206 / L282-L298: wrap along the axis (period n - 1 intervals), then
        repeat across it."""
        n, nb = self._n(), self.nb
        st_a, st_o = self._staggered(field_name)
        mx, my = n + int(st_a), 1 + int(st_o)
        mi, mid = mx + 2 * nb, slice(nb, my + nb)
        field[self._ix(slice(0, nb), mid)] = field[self._ix(slice(n - 1, n - 1 + nb), mid)]
        lo = nb + 1 if mx == n else nb + 2
        field[self._ix(slice(mx + nb, mx + 2 * nb), mid)] = field[self._ix(slice(lo, lo + nb), mid)]
        al = slice(0, mi)
        field[self._ix(al, slice(0, nb))] = field[self._ix(al, slice(nb, nb + 1))]
        src = slice(nb, nb + 1) if my == 1 else slice(nb + 1, nb + 2)
        field[self._ix(al, slice(my + nb, my + 2 * nb))] = field[self._ix(al, src)]

    def get_numerical_field(self, field, field_name=None):
        """periodic.py:L156-L185 / L248-L277"""
        n, nb = self._n(), self.nb
        st_a, st_o = self._staggered(field_name)
        mx, my = n + int(st_a), 1 + int(st_o)
        src = storage.as_storage(field, device=self.storage_options.device)
        trg = self.zeros(shape=(src.shape[0] + 2 * nb, src.shape[1] + 2 * nb) + tuple(src.shape[2:]))
        trg[self._ix(slice(nb, mx + nb), slice(nb, my + nb))] = src[self._ix(slice(0, mx), slice(0, my))]
        self.enforce_field(trg, field_name)
        return trg

    def get_physical_field(self, field, field_name=None):
        return field[self.nb : -self.nb, self.nb : -self.nb]

    def set_outermost_layers_x(self, field, field_name=None, field_units=None, time=None):
        field[0, :] = field[-2, :]
        field[-1, :] = field[1, :]

    def set_outermost_layers_y(self, field, field_name=None, field_units=None, time=None):
        field[:, 0] = field[:, -2]
        field[:, -1] = field[:, 1]


class Periodic1DX(_Periodic1D):
    """periodic.py:L125-L214 (ny == 1)."""

    axis = 0


class Periodic1DY(_Periodic1D):
    """periodic.py:L217-L306 (nx == 1)."""

    axis = 1
