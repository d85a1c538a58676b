# -*- coding: utf-8 -*-
"""The reusable numerics ("dwarfs") of the hot path on b200 storages, with the reference's
class and method names:

  HorizontalDiffusion (second_order, fourth_order)  src/tasmania/dwarfs/horizontal_diffusion.py:L41-L175
  HorizontalSmoothing (first..third_order)          src/tasmania/dwarfs/horizontal_smoothing.py:L41-L134
  (both with the one-dimensional ``..._1dx`` / ``..._1dy`` variants of the reference's
  subclasses/horizontal_diffusers/*.py and subclasses/horizontal_smoothers/*.py)
  VerticalDamping (rayleigh)                        src/tasmania/dwarfs/vertical_damping.py:L46-L175
  HorizontalVelocity, WaterConstituent              src/tasmania/dwarfs/diagnostics.py:L44-L466

Coefficient profiles are rank-1 in k by construction in the reference; here they are stored
once as ``(1, 1, nk)`` and handed to the kernels through zero-stride views instead of being
materialised (and streamed from HBM) as full 3-D storages.
"""
from __future__ import annotations

import math

import numpy as np

from tasmania_b200 import storage
from tasmania_b200.framework import BackendOptions, GridComponent, StencilFactory, StorageOptions


def _profile_storage(profile, shape, device):
    """(1, 1, nk) storage + its (ni, nj, nk) zero-stride broadcast view."""
    p = storage.as_storage(np.asarray(profile, dtype=float)[None, None, :], device=device)
    return p, storage.B200Array(p.t.expand(shape[0], shape[1], -1))


def vertical_profile(coeff, coeff_max, damp_depth, nk):
    """gamma(k), horizontal_diffusion.py:L91-L97 / horizontal_smoothing.py:L83-L89."""
    gamma = coeff * np.ones(nk)
    n = damp_depth
    if n > 0:
        pert = np.sin(0.5 * math.pi * (n - np.arange(0, n, dtype=float)) / n) ** 2
        gamma[:n] += (coeff_max - coeff) * pert
    return gamma


def _one_dimensional(orders):
    """name -> (order, axis): the 2-D schemes (axis None) and their _1dx / _1dy variants"""
    out = {name: (order, None) for name, order in orders.items()}
    for name, order in orders.items():
        out[name + "_1dx"], out[name + "_1dy"] = (order, 0), (order, 1)
    return out


def _interior(shape, nb, axis):
    """origin and domain of the interior: nb points off both ends of the stencil axis (or axes)"""
    nx, ny, nz = shape
    bx, by = (nb if axis in (None, 0) else 0), (nb if axis in (None, 1) else 0)
    return (bx, by, 0), (nx - 2 * bx, ny - 2 * by, nz)


class HorizontalDiffusion(StencilFactory):
    """Tendency due to horizontal diffusion; ``factory("second_order" | "fourth_order" |
    "second_order_1dx" | ... | "fourth_order_1dy", ...)``."""

    ORDERS = _one_dimensional({"second_order": 2, "fourth_order": 4})

    def __init__(self, diffusion_type, shape, dx, dy, diffusion_coeff, diffusion_coeff_max,
                 diffusion_damp_depth, nb=None, *, backend="b200", backend_options=None,
                 storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        if diffusion_type not in self.ORDERS:
            raise ValueError(f"unknown (or out-of-scope) diffusion type {diffusion_type!r}")
        self.order, self.axis = self.ORDERS[diffusion_type]
        min_nb = self.order // 2
        nb = min_nb if (nb is None or nb < min_nb) else nb
        lb = 2 * nb + 1
        assert all(shape[a] >= lb for a in (0, 1) if self.axis in (None, a))
        self._shape, self._nb, self._dx, self._dy = tuple(shape), nb, dx, dy
        gamma = vertical_profile(diffusion_coeff, diffusion_coeff_max, diffusion_damp_depth, shape[2])
        self._gamma1d, self._gamma = _profile_storage(gamma, shape, self.storage_options.device)
        self.backend_options.externals = {
            "set_output": self.get_subroutine_definition("set_output"),
            "diffusion_order": self.order,
            "diffusion_axis": self.axis,
        }
        self._stencil = self.compile_stencil("diffusion")

    @classmethod
    def factory(cls, diffusion_type, *args, **kwargs):
        return cls(diffusion_type, *args, **kwargs)

    def __call__(self, phi, phi_tnd, *, overwrite_output=True):
        origin, domain = _interior(self._shape, self._nb, self.axis)
        self._stencil(in_phi=phi, in_gamma=self._gamma, out_phi=phi_tnd, dx=self._dx, dy=self._dy,
                      ow_out_phi=overwrite_output, origin=origin, domain=domain)


class HorizontalSmoothing(StencilFactory):
    """Horizontal numerical smoothing; ``factory("first_order" | ... | "third_order" |
    "first_order_1dx" | ... | "third_order_1dy", ...)``.
    The reference's five launches (smoothing + four rim copies; three for a 1-D smoother) are
    one kernel here."""

    ORDERS = _one_dimensional({"first_order": 1, "second_order": 2, "third_order": 3})

    def __init__(self, smooth_type, shape, smooth_coeff, smooth_coeff_max, smooth_damp_depth,
                 nb=None, *, backend="b200", backend_options=None, storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        if smooth_type not in self.ORDERS:
            raise ValueError(f"unknown (or out-of-scope) smoothing type {smooth_type!r}")
        self.order, self.axis = self.ORDERS[smooth_type]
        nb = self.order if (nb is None or nb < self.order) else nb
        lb = 2 * nb + 1
        assert all(shape[a] >= lb for a in (0, 1) if self.axis in (None, a))
        self._shape, self._nb = tuple(shape), nb
        gamma = vertical_profile(smooth_coeff, smooth_coeff_max, smooth_damp_depth, shape[2])
        self._gamma1d, self._gamma = _profile_storage(gamma, shape, self.storage_options.device)
        self.backend_options.externals = {"smoothing_order": self.order, "rim_copy": True,
                                          "smoothing_axis": self.axis}
        self._stencil_smooth = self.compile_stencil("smoothing")

    @classmethod
    def factory(cls, smooth_type, *args, **kwargs):
        return cls(smooth_type, *args, **kwargs)

    def __call__(self, phi, phi_out):
        origin, domain = _interior(self._shape, self._nb, self.axis)
        self._stencil_smooth(in_phi=phi, in_gamma=self._gamma, out_phi=phi_out,
                             origin=origin, domain=domain)


def rayleigh_coefficient(z_main, z_top, damp_depth, damp_max, nk):
    """vertical_damping.py:L100-L111."""
    nz = len(z_main)
    r = np.zeros(nk)
    if damp_depth > 0:
        z = np.concatenate((z_main, np.array([0]))) if nk == nz + 1 else np.asarray(z_main)
        za = z[damp_depth - 1]
        r = (z >= za) * damp_max * (1 - np.cos(math.pi * (z - za) / (z_top - za)))
    return r


class VerticalDamping(GridComponent, StencilFactory):
    """Rayleigh wave absorber: ``out = new - dt R(k) (now - ref)``."""

    def __init__(self, damp_type, grid, damp_depth=15, damp_coeff_max=0.0002, time_units="s", *,
                 backend="b200", backend_options=None, storage_shape=None, storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        if damp_type != "rayleigh":
            raise ValueError(f"unknown vertical damping type {damp_type!r}")
        assert damp_depth <= grid.nz
        self._grid = grid
        self._damp_depth = damp_depth
        self._shape = tuple(storage_shape or (grid.nx + 1, grid.ny + 1, grid.nz + 1))
        r = rayleigh_coefficient(grid.z, grid.z_on_interface_levels[0], damp_depth, damp_coeff_max,
                                 self._shape[2])
        self._rmat1d, self._rmat = _profile_storage(r, self._shape, self.storage_options.device)
        self._stencil_damp = self.compile_stencil("damping")

    @classmethod
    def factory(cls, damp_type, *args, **kwargs):
        return cls(damp_type, *args, **kwargs)

    def __call__(self, dt, field_now, field_new, field_ref, field_out):
        self._stencil_damp(in_phi_now=field_now, in_phi_new=field_new, in_phi_ref=field_ref,
                           in_rmat=self._rmat, out_phi=field_out, dt=dt.total_seconds(),
                           origin=(0, 0, 0), domain=self._shape)


class HorizontalVelocity(GridComponent, StencilFactory):
    """Momenta <-> velocity components, dwarfs/diagnostics.py:L44-L272."""

    def __init__(self, grid, staggering=True, *, backend="b200", backend_options=None,
                 storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        self._grid, self._staggering = grid, staggering
        self.backend_options.externals = {"staggering": staggering}
        self._stencil_diagnosing_momenta = self.compile_stencil("momenta")
        self._stencil_diagnosing_velocity_x = self.compile_stencil("velocity_x")
        self._stencil_diagnosing_velocity_y = self.compile_stencil("velocity_y")

    def get_momenta(self, d, u, v, du, dv):
        g = self._grid
        self._stencil_diagnosing_momenta(in_d=d, in_u=u, in_v=v, out_du=du, out_dv=dv,
                                         origin=(0, 0, 0), domain=(g.nx, g.ny, g.nz))

    def get_velocity_components(self, d, du, dv, u, v):
        g, dn = self._grid, int(self._staggering)
        self._stencil_diagnosing_velocity_x(in_d=d, in_du=du, out_u=u, origin=(dn, 0, 0),
                                            domain=(g.nx - dn, g.ny, g.nz))
        self._stencil_diagnosing_velocity_y(in_d=d, in_dv=dv, out_v=v, origin=(0, dn, 0),
                                            domain=(g.nx, g.ny - dn, g.nz))


class WaterConstituent(GridComponent, StencilFactory):
    """Density <-> mass fraction of a water species, dwarfs/diagnostics.py:L275-L466."""

    def __init__(self, grid, clipping=False, *, backend="b200", backend_options=None,
                 storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        self._grid = grid
        self.backend_options.externals = {"clipping": clipping}
        self._stencil_diagnosing_density = self.compile_stencil("density")
        self._stencil_diagnosing_mass_fraction = self.compile_stencil("mass_fraction")

    def get_density_of_water_constituent(self, d, q, dq):
        g = self._grid
        self._stencil_diagnosing_density(in_d=d, in_q=q, out_dq=dq, origin=(0, 0, 0),
                                         domain=(g.nx, g.ny, g.nz))

    def get_mass_fraction_of_water_constituent_in_air(self, d, dq, q):
        g = self._grid
        self._stencil_diagnosing_mass_fraction(in_d=d, in_dq=dq, out_q=q, origin=(0, 0, 0),
                                               domain=(g.nx, g.ny, g.nz))
