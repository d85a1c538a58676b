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
    # periodic fixtures store the sizes of the numerical grid (physical + nb ghost points a side)
    periodic = "boundary" in fx.files and str(fx["boundary"][0]) == "periodic"
    hb = ob.Periodic(nx - 2 * nb, ny - 2 * nb, nz, nb) if periodic else ob.Relaxed(nx, ny, nz, nb, nr)
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


def moist_case(nx, ny, nz, **kwargs):
    """The moist mountain-flow case (BASELINE config 3, scalable): see
    tasmania_b200.isentropic_moist.moist_mountain_case, which builds it (bench.py uses it too)."""
    from tasmania_b200.isentropic_moist import moist_mountain_case

    return moist_mountain_case(nx, ny, nz, **kwargs)


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
