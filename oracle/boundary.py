# -*- coding: utf-8 -*-
"""Oracle (test infrastructure): lateral boundary conditions, rows K5 of SURVEY.md section 8a.

Follows
  src/tasmania/framework/subclasses/stencil_definitions/algorithms.py:L32-L57  (irelax, relax)
  src/tasmania/domain/subclasses/horizontal_boundaries/relaxed.py:L119-L247    (Relaxed)
  src/tasmania/domain/subclasses/horizontal_boundaries/periodic.py:L64-L122    (Periodic)
  src/tasmania/domain/horizontal_boundary.py:L299-L344                         (enforce_raw)
"""
import numpy as np


# ------------------------------------------------------------------ stencils
def irelax(gamma, phi_ref, phi, origin, domain):
    """algorithms.py:L32-L43 -- in place; gamma==0 keeps, gamma==1 takes the reference."""
    idx = tuple(slice(o, o + d) for o, d in zip(origin, domain))
    g, r, p = gamma[idx], phi_ref[idx], phi[idx]
    phi[idx] = np.where(g == 0.0, p, np.where(g == 1.0, r, p - g * (p - r)))


def relax(gamma, phi, phi_ref, out, origin, domain):
    """algorithms.py:L46-L57."""
    idx = tuple(slice(o, o + d) for o, d in zip(origin, domain))
    g, r, p = gamma[idx], phi_ref[idx], phi[idx]
    out[idx] = np.where(g == 0.0, p, np.where(g == 1.0, r, p - g * (p - r)))


def _stagger(nx, ny, nz, name):
    """Extent of the computational domain of a field, relaxed.py:L124-L139."""
    name = name or ""
    mi = nx + 1 if ("at_u_locations" in name or "at_uv_locations" in name) else nx
    mj = ny + 1 if ("at_v_locations" in name or "at_uv_locations" in name) else ny
    mk = nz + 1 if "on_interface_levels" in name else nz
    return mi, mj, mk


# ------------------------------------------------------------------ Relaxed
def relaxed_gamma(nx, ny, nz, nb, nr, shape=None):
    """Coefficient matrix of the relaxed boundary, relaxed.py:L193-L247."""
    rel = np.array([1.0] + [1.0 - np.tanh(0.5 * m) for m in range(1, 8)])
    rel = rel[:nr]
    rel[:nb] = 1.0
    rrel = rel[::-1]

    shape = shape or (nx + 1, ny + 1, nz + 1)
    g = np.zeros(shape)

    corner = np.zeros((nr, nr))
    for i in range(nr):
        corner[i, i:] = rel[i]
        corner[i:, i] = rel[i]
    xpyn = corner[::-1, :]
    xpyp = xpyn[:, ::-1]
    xnyp = corner[:, ::-1]

    g[:nr, :nr] = corner[:, :, None]
    g[:nr, nr : ny - nr] = rel[:, None, None]
    g[:nr, ny - nr : ny] = xnyp[:, :, None]
    g[nx - nr : nx, :nr] = xpyn[:, :, None]
    g[nx - nr : nx, nr : ny - nr] = rrel[:, None, None]
    g[nx - nr : nx, ny - nr : ny] = xpyp[:, :, None]
    g[nr : nx - nr, :nr] = rel[None, :, None]
    g[nr : nx - nr, ny - nr : ny] = rrel[None, :, None]
    g[nx : nx + 1, : ny + 1] = 1.0
    g[: nx + 1, ny : ny + 1] = 1.0
    return g


class Relaxed:
    """State-less restatement of ``Relaxed`` acting on raw arrays."""

    type = "relaxed"

    def __init__(self, nx, ny, nz, nb, nr=8, shape=None):
        assert nb <= nr <= 8 and nr <= nx / 2 and nr <= ny / 2
        self.nx, self.ny, self.nz, self.nb, self.nr = nx, ny, nz, nb, nr
        self.ni, self.nj = nx, ny
        self.gamma = relaxed_gamma(nx, ny, nz, nb, nr, shape)
        self.reference_state = {}

    def enforce_field(self, field, field_name=None):
        mi, mj, mk = _stagger(self.nx, self.ny, self.nz, field_name)
        irelax(self.gamma, self.reference_state[field_name], field, (0, 0, 0), (mi, mj, mk))

    def enforce_raw(self, state, field_names=None):
        """horizontal_boundary.py:L299-L344 -- every field that has a reference value."""
        for name in state:
            if name == "time" or name not in self.reference_state:
                continue
            if field_names is not None and name not in field_names:
                continue
            self.enforce_field(state[name], name)

    def set_outermost_layers_x(self, field, field_name=None):
        mi, mj, _ = _stagger(self.nx, self.ny, self.nz, field_name)
        ref = self.reference_state[field_name]
        field[0, :mj] = ref[0, :mj]
        field[mi - 1, :mj] = ref[mi - 1, :mj]

    def set_outermost_layers_y(self, field, field_name=None):
        mi, mj, _ = _stagger(self.nx, self.ny, self.nz, field_name)
        ref = self.reference_state[field_name]
        field[:mi, 0] = ref[:mi, 0]
        field[:mi, mj - 1] = ref[:mi, mj - 1]


# ------------------------------------------------------------------ Periodic
class Periodic:
    """Periodic conditions on the numerical grid (physical grid + nb ghost points a side).

    ``nx, ny`` are the *physical* sizes; fields live on ``(nx + 2 nb [+1], ny + 2 nb [+1])``.
    """

    type = "periodic"

    def __init__(self, nx, ny, nz, nb):
        assert nb <= nx / 2 and nb <= ny / 2
        self.nx, self.ny, self.nz, self.nb = nx, ny, nz, nb
        self.ni, self.nj = nx + 2 * nb, ny + 2 * nb
        self.reference_state = {}

    def get_numerical_field(self, field, field_name=None):
        """periodic.py:L64-L96."""
        nx, ny, nb = self.nx, self.ny, self.nb
        mx, my, _ = _stagger(nx, ny, self.nz, field_name)
        shape = (field.shape[0] + 2 * nb, field.shape[1] + 2 * nb) + tuple(field.shape[2:])
        trg = np.zeros(shape, dtype=field.dtype)
        trg[nb : mx + nb, nb : my + nb] = field[:mx, :my]
        self.enforce_field(trg, field_name)
        return trg

    def get_physical_field(self, field, field_name=None):
        return field[self.nb : -self.nb, self.nb : -self.nb]

    def enforce_field(self, field, field_name=None):
        """periodic.py:L98-L122 -- x first, then y over the already extended i-range."""
        nx, ny, nb = self.nx, self.ny, self.nb
        mx, my, _ = _stagger(nx, ny, self.nz, field_name)
        mi = mx + 2 * nb
        sx = 1 if mx == nx else 2
        sy = 1 if my == ny else 2
        field[:nb, nb : my + nb] = field[nx - 1 : nx - 1 + nb, nb : my + nb]
        field[mx + nb : mx + 2 * nb, nb : my + nb] = field[nb + sx : 2 * nb + sx, nb : my + nb]
        field[:mi, :nb] = field[:mi, ny - 1 : ny - 1 + nb]
        field[:mi, my + nb : my + 2 * nb] = field[:mi, nb + sy : 2 * nb + sy]

    def enforce_raw(self, state, field_names=None):
        # base-class quirk (horizontal_boundary.py:L322-L334): only fields that appear in
        # the reference state are touched, whatever the boundary type
        for name in state:
            if name == "time" or name not in self.reference_state:
                continue
            if field_names is not None and name not in field_names:
                continue
            self.enforce_field(state[name], name)

    def set_outermost_layers_x(self, field, field_name=None):
        field[0, :] = field[-2, :]
        field[-1, :] = field[1, :]

    def set_outermost_layers_y(self, field, field_name=None):
        field[:, 0] = field[:, -2]
        field[:, -1] = field[:, 1]
