# -*- coding: utf-8 -*-
"""Shared test helpers: build oracle objects out of a golden fixture."""
import os
from datetime import datetime, timedelta

import numpy as np

from oracle import boundary as ob
from oracle import isentropic as oi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
S, SU, SV, U, V, MTG = oi.S, oi.SU, oi.SV, oi.U, oi.V, oi.MTG
P = "air_pressure_on_interface_levels"
EXN = "exner_function_on_interface_levels"
H = "height_on_interface_levels"


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


class Topography:
    """Time-growing topography, src/tasmania/domain/topography.py:L75-L81, L106-L116."""

    def __init__(self, steady, grow_seconds):
        self.steady = np.asarray(steady)
        self.time = timedelta(seconds=grow_seconds)
        self.fact = float(self.time.total_seconds() == 0.0)
        self.profile = self.fact * self.steady

    def update(self, elapsed):
        if self.fact < 1.0:
            self.fact = min(elapsed / self.time, 1.0)
            self.profile = self.fact * self.steady

    def __call__(self):
        return self.profile


def relerr(a, b):
    """Norm-wise relative error max|a-b| / max|b| (0 if both vanish)."""
    scale = float(np.max(np.abs(b)))
    diff = float(np.max(np.abs(np.asarray(a) - np.asarray(b))))
    return diff / scale if scale > 0 else diff


def oracle_dry_run(fx, nsteps=None):
    """Run the oracle dycore on a ``isen_dry_*`` fixture; returns (final dict, stage0 dict)."""
    nx, ny, nz, nb, nr, nsteps_fx, damp_depth, damp_every = (int(v) for v in fx["dims"])
    nsteps = nsteps or nsteps_fx
    dx, dy, dz, pt, dt_s, eps, damp_max, topo_time = (float(v) for v in fx["params"])
    scheme, flux = (str(v) for v in fx["scheme"])
    grid = oi.Grid(nx, ny, nz, dx, dy, dz, fx["z_hl"], fx["z"])
    hb = ob.Relaxed(nx, ny, nz, nb, nr)
    moist = ("init_" + oi.MFWV) in fx.files
    qnames = (oi.MFWV, oi.MFCW, oi.MFPW) if moist else ()
    names = (S, MTG, SU, U, SV, V, P, EXN, H) + qnames
    state = {n: fx["init_" + n].copy() for n in names}
    state["time"] = datetime(2000, 1, 1)
    hb.reference_state = {n: state[n].copy() for n in names}
    topo = Topography(fx["topo_steady"], topo_time)
    dyc = oi.IsentropicDycore(grid, hb, topo, moist=moist, scheme=scheme, flux=flux, pt=pt, eps=eps,
                              damp=True, damp_at_every_stage=bool(damp_every),
                              damp_depth=damp_depth, damp_max=damp_max)
    dt = timedelta(seconds=dt_s)
    stage0 = None
    for step in range(nsteps):
        topo.update((step + 1) * dt)
        out = dyc(state, {}, dt)
        if step == 0:
            stage0 = {n: dyc._stage_states[0][n].copy() if dyc.stages > 1 else out[n].copy()
                      for n in (S, SU, U, SV, V) + qnames}
        new = {n: out[n].copy() for n in (S, SU, U, SV, V) + qnames}
        new["time"] = out["time"]
        for n in (P, EXN, H, MTG):
            new[n] = state[n].copy()
        oi.refresh_diagnostics(grid, topo(), new[S], pt, new[P], new[EXN], new[MTG], new[H])
        state = new
    return state, stage0, (grid, hb, topo, dyc)


def moist_case(nx, ny, nz, *, topo_seconds=60.0, max_height=1000.0, relative_humidity=0.98,
               seed=True, half_width_km=(176.0, 176.0)):
    """BASELINE config 3 scaled down: the moist mountain-flow case of namelist_sus.py on a 352 km
    square (grid spacing of a 161-point axis is kept when nx = ny = 161), with a faster-growing,
    taller mountain and -- when ``seed`` -- blobs of cloud water and rain in the initial state so
    that autoconversion, accretion, evaporation, sedimentation and precipitation are all active
    within a few steps.  Returns (tasmania_b200 Grid, numpy state)."""
    from tasmania_b200.grid import Grid, Topography as GTopography
    from tasmania_b200.grid import gaussian_profile, isentropic_state_from_brunt_vaisala

    hx, hy = half_width_km
    x = np.linspace(-hx, hx, nx)
    y = np.linspace(-hy, hy, ny)
    topo = GTopography(gaussian_profile(x, y, max_height, 50.0, 50.0), timedelta(seconds=topo_seconds))
    grid = Grid((-hx, hx), nx, (-hy, hy), ny, (400.0, 280.0), nz, units_to_m=1e3,
                topography=topo)
    state = isentropic_state_from_brunt_vaisala(grid, 22.5, 0.0, 0.015, moist=True,
                                                precipitation=True,
                                                relative_humidity=relative_humidity)
    if seed:
        i, j, k = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
        blob = np.exp(-(((i - nx // 2) / (0.2 * nx)) ** 2 + ((j - ny // 2) / (0.2 * ny)) ** 2
                        + ((k - 0.7 * nz) / (0.2 * nz)) ** 2))
        blob[nx:, :, :] = blob[:, ny:, :] = blob[:, :, nz:] = 0.0
        state[oi.MFCW] = 8e-4 * blob
        state[oi.MFPW] = 3e-4 * np.roll(blob, 2, axis=0) * (blob > 0)
    return grid, state


BOUNDARY_1D_NAMES = ("air_isentropic_density", "x_velocity_at_u_locations", "y_velocity_at_v_locations",
                     "air_pressure_on_interface_levels")


def check_one_dimensional_boundaries(fx, tb):
    """The b200 mirrors of Relaxed1DX / 1DY and Periodic1DX / 1DY against what the reference's own
    classes wrote into tests/golden/stencils_1d.npz (shared by the host test, which runs it through
    the oracle-backed ABI stub, and the GPU test)."""
    from tasmania_b200.boundary import HorizontalBoundary

    for kind in ("relaxed", "periodic"):
        for ax in "xy":
            tag = f"hb_{kind}_{ax}"
            nx, ny, nz, nb, nr = (int(v) for v in fx[tag + "_dims"])
            hb = HorizontalBoundary.factory(kind, nx, ny, nz, nb, **({"nr": nr} if kind == "relaxed" else {}))
            num = hb.get_numerical_field(tb.as_storage(fx[tag + "_phys"]), field_name=BOUNDARY_1D_NAMES[0])
            np.testing.assert_array_equal(tb.to_numpy(num), fx[tag + "_num"], err_msg=tag)
            np.testing.assert_array_equal(tb.to_numpy(hb.get_physical_field(num)), fx[tag + "_phys"])
            hb.reference_state = {n: tb.as_storage(fx[f"{tag}_ref{m}"])
                                  for m, n in enumerate(BOUNDARY_1D_NAMES)}
            for m, n in enumerate(BOUNDARY_1D_NAMES):
                f = tb.as_storage(fx[f"{tag}_in{m}"])
                hb.enforce_field(f, field_name=n)
                hb.set_outermost_layers_x(f, field_name=n)
                hb.set_outermost_layers_y(f, field_name=n)
                np.testing.assert_array_equal(tb.to_numpy(f), fx[f"{tag}_out{m}"], err_msg=f"{tag} {n}")
