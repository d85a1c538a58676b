# -*- coding: utf-8 -*-
"""Host-side set-up objects: grid, topography, initial isentropic state.

Out of the hot path by SURVEY.md section 2 (rows 6 and 10: "host-side set-up, O(nx ny) numpy"),
but needed on the GPU box to build benchmark inputs, where the reference is not available.
Plain numpy, evaluated once; formulas follow
  src/tasmania/domain/horizontal_grid.py:L80-L120, src/tasmania/domain/grid.py:L278-L312
  src/tasmania/domain/topography.py:L75-L116, subclasses/topographies/gaussian.py:L100-L135
  src/tasmania/isentropic/state.py:L126-L230
and are checked against reference-generated fixtures in tests/test_host_setup.py.
"""
from __future__ import annotations

import math
from datetime import timedelta
from typing import Optional

import numpy as np

# default physical constants of the isentropic model
# (src/tasmania/isentropic/dynamics/diagnostics.py:L56-L63, isentropic/state.py:L44-L58)
CONSTANTS = {"pref": 1.0e5, "rd": 287.05, "g": 9.80665, "cp": 1004.0}
# the state builder has its own defaults -- note g (src/tasmania/isentropic/state.py:L45-L52)
STATE_CONSTANTS = {"pref": 1.0e5, "rd": 287.05, "g": 9.81, "cp": 1004.0}


class Topography:
    """Time-growing terrain height: ``profile = min(t / t_grow, 1) * steady_profile``."""

    def __init__(self, steady_profile: np.ndarray, grow_time: Optional[timedelta] = None):
        self.steady_profile = np.asarray(steady_profile, dtype=float)
        self.time = grow_time or timedelta(seconds=0)
        self._fact = float(self.time.total_seconds() == 0.0)
        self._profile, self._profile_fact = None, None

    @property
    def profile(self) -> np.ndarray:
        """Host copy of the current height, evaluated lazily (the device path only needs
        the growth factor, see IsentropicDiagnostics._set_topography)."""
        if self._profile_fact != self._fact:
            self._profile = self._fact * self.steady_profile
            self._profile_fact = self._fact
        return self._profile

    def update(self, elapsed: timedelta) -> None:
        if self._fact < 1.0:
            self._fact = min(elapsed / self.time, 1.0)


def gaussian_profile(x, y, max_height=500.0, width_x=1.0, width_y=1.0, center_x=None, center_y=None):
    """gaussian.py:L100-L135; x, y, widths and centres in the same (native) units, height in m."""
    cx = 0.5 * (x[0] + x[-1]) if center_x is None else center_x
    cy = 0.5 * (y[0] + y[-1]) if center_y is None else center_y
    xx, yy = np.meshgrid(x, y, indexing="ij")
    return max_height * np.exp(-(((xx - cx) / width_x) ** 2) - ((yy - cy) / width_y) ** 2)


class Grid:
    """Three-dimensional grid: regular in x, y; isentropic (theta) levels in z, top first.

    ``domain_x`` / ``domain_y`` are given in native units (e.g. km) with ``units_to_m`` the
    conversion factor, as the reference builds the axes in native units and converts
    afterwards (the order matters in the last bit).
    """

    def __init__(self, domain_x, nx, domain_y, ny, domain_z, nz, *, units_to_m=1.0,
                 topography: Optional[Topography] = None, x=None, y=None):
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.x_native = np.linspace(domain_x[0], domain_x[1], nx) if x is None else np.asarray(x)
        self.y_native = np.linspace(domain_y[0], domain_y[1], ny) if y is None else np.asarray(y)
        dx = 1.0 if nx == 1 else (self.x_native[-1] - self.x_native[0]) / (nx - 1)
        dy = 1.0 if ny == 1 else (self.y_native[-1] - self.y_native[0]) / (ny - 1)
        self.dx_native = dx if dx != 0.0 else 1.0
        self.dy_native = dy if dy != 0.0 else 1.0
        self.units_to_m = units_to_m
        self.x = self.x_native * units_to_m
        self.y = self.y_native * units_to_m
        self.dx = self.dx_native * units_to_m
        self.dy = self.dy_native * units_to_m
        self.z_on_interface_levels = np.linspace(domain_z[0], domain_z[1], nz + 1)
        self.z = 0.5 * (self.z_on_interface_levels[:-1] + self.z_on_interface_levels[1:])
        dz = math.fabs(self.z_on_interface_levels[0] - self.z_on_interface_levels[-1]) / nz
        self.dz = 1.0 if dz == 0.0 else dz
        self.topography = topography or Topography(np.zeros((nx, ny)))

    def update_topography(self, elapsed: timedelta) -> None:
        self.topography.update(elapsed)

    def extended(self, nb: int) -> "Grid":
        """Numerical grid of a periodic domain: nb ghost points a side
        (src/tasmania/domain/subclasses/horizontal_boundaries/periodic.py:L44-L50)."""
        xn = np.concatenate((self.x_native[0] + self.dx_native * np.arange(-nb, 0), self.x_native,
                             self.x_native[-1] + self.dx_native * np.arange(1, nb + 1)))
        yn = np.concatenate((self.y_native[0] + self.dy_native * np.arange(-nb, 0), self.y_native,
                             self.y_native[-1] + self.dy_native * np.arange(1, nb + 1)))
        g = Grid((xn[0], xn[-1]), self.nx + 2 * nb, (yn[0], yn[-1]), self.ny + 2 * nb,
                 (self.z_on_interface_levels[0], self.z_on_interface_levels[-1]), self.nz,
                 units_to_m=self.units_to_m, x=xn, y=yn)
        g.dx_native, g.dy_native = self.dx_native, self.dy_native
        g.dx, g.dy = self.dx, self.dy
        return g


def isentropic_state_from_brunt_vaisala(grid: Grid, x_velocity: float, y_velocity: float,
                                        brunt_vaisala: float, constants=STATE_CONSTANTS,
                                        storage_shape=None, moist=False, precipitation=False,
                                        relative_humidity=0.5) -> dict:
    """Initial state (numpy arrays), src/tasmania/isentropic/state.py:L126-L391; with ``moist``
    also air density / temperature, the water vapour of a uniform relative humidity (Tetens,
    src/tasmania/utils/meteo.py:L192-L274), no cloud water, no rain."""
    nx, ny, nz = grid.nx, grid.ny, grid.nz
    dz, hs, bv = grid.dz, grid.topography.profile, brunt_vaisala
    rd, g, pref, cp = constants["rd"], constants["g"], constants["pref"], constants["cp"]
    shape = tuple(storage_shape or (nx + 1, ny + 1, nz + 1))

    u = np.zeros(shape)
    u[: nx + 1, :ny, :nz] = x_velocity
    v = np.zeros(shape)
    v[:nx, : ny + 1, :nz] = y_velocity

    theta1d = grid.z[np.newaxis, np.newaxis, :]
    h = np.zeros(shape)
    h[:nx, :ny, nz] = hs
    for k in range(nz - 1, -1, -1):
        h[:nx, :ny, k : k + 1] = h[:nx, :ny, k + 1 : k + 2] + g * dz / (
            (bv**2) * theta1d[:, :, k : k + 1]
        )
    exn = np.zeros(shape)
    exn[:nx, :ny, nz] = cp
    for k in range(nz - 1, -1, -1):
        exn[:nx, :ny, k : k + 1] = exn[:nx, :ny, k + 1 : k + 2] - dz * (g**2) / (
            (bv**2) * (theta1d[:, :, k : k + 1] ** 2)
        )
    p = np.zeros(shape)
    p[:nx, :ny, : nz + 1] = pref * ((exn[:nx, :ny, : nz + 1] / cp) ** (cp / rd))

    mtg_s = g * h[:, :, nz : nz + 1] + grid.z_on_interface_levels[-1] * exn[:, :, nz : nz + 1]
    mtg = np.zeros(shape)
    mtg[:nx, :ny, nz - 1] = mtg_s[:nx, :ny, 0] + 0.5 * dz * exn[:nx, :ny, nz]
    for k in range(nz - 2, -1, -1):
        mtg[:nx, :ny, k] = mtg[:nx, :ny, k + 1] + dz * exn[:nx, :ny, k + 1]

    s = np.zeros(shape)
    s[:nx, :ny, :nz] = -(p[:nx, :ny, :nz] - p[:nx, :ny, 1 : nz + 1]) / (g * dz)
    su = np.zeros(shape)
    su[:nx, :ny, :nz] = 0.5 * s[:nx, :ny, :nz] * (u[:nx, :ny, :nz] + u[1 : nx + 1, :ny, :nz])
    sv = np.zeros(shape)
    sv[:nx, :ny, :nz] = 0.5 * s[:nx, :ny, :nz] * (v[:nx, :ny, :nz] + v[:nx, 1 : ny + 1, :nz])

    state = {
        "air_isentropic_density": s,
        "air_pressure_on_interface_levels": p,
        "exner_function_on_interface_levels": exn,
        "height_on_interface_levels": h,
        "montgomery_potential": mtg,
        "x_momentum_isentropic": su,
        "x_velocity_at_u_locations": u,
        "y_momentum_isentropic": sv,
        "y_velocity_at_v_locations": v,
    }
    if moist:  # state.py:L286-L389
        rho = np.zeros(shape)
        rho[:nx, :ny, :nz] = s[:nx, :ny, :nz] * dz / (h[:nx, :ny, :nz] - h[:nx, :ny, 1 : nz + 1])
        temp = np.zeros(shape)
        temp[:nx, :ny, :nz] = (
            0.5 * (exn[:nx, :ny, :nz] + exn[:nx, :ny, 1 : nz + 1]) * theta1d[:, :, :nz] / cp
        )
        p_unstg = np.zeros(shape)
        p_unstg[:nx, :ny, :nz] = 0.5 * (p[:nx, :ny, :nz] + p[:nx, :ny, 1 : nz + 1])
        rh = np.full(shape, float(relative_humidity))
        # meteo.py:L231-L248, L266-L272
        p_sat = 610.78 * np.exp(17.27 * (temp - 273.16) / (temp - 35.86))
        pw = rh * p_sat
        with np.errstate(divide="ignore", invalid="ignore"):
            qv = np.where(p_sat >= 0.616 * p_unstg, 0, 0.62198 * pw / (p_unstg - pw))
        state["air_density"], state["air_temperature"] = rho, temp
        state["mass_fraction_of_water_vapor_in_air"] = qv
        state["mass_fraction_of_cloud_liquid_water_in_air"] = np.zeros(shape)
        state["mass_fraction_of_precipitation_water_in_air"] = np.zeros(shape)
        if precipitation:
            state["precipitation"] = np.zeros((shape[0], shape[1], 1))
            state["accumulated_precipitation"] = np.zeros((shape[0], shape[1], 1))
    return state
