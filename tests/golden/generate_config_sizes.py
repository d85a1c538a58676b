# -*- coding: utf-8 -*-
"""Fixtures at the grid sizes and step counts BASELINE.json / north_star name, produced by the
numpy ORACLE (oracle/, itself held bit for bit to the reference's own code by
tests/test_oracle_golden.py and the *_reference tests):

  config_c2.npz   dry isentropic, 161 x 161 x 60, RK3WS + fifth-order upwind, 100 steps
  config_c3.npz   moist isentropic + Kessler + sedimentation (SUS), 256 x 256 x 60, 20 steps
  config_c5k.npz  dry isentropic, 512 x 512 x 64 (the kernel selection of config 5), 3 steps

The runs take minutes of host time each (C3: 23 s per step on one core), too long to repeat
inside a GPU test, and the full final states are 0.1 - 0.6 GB; the fixtures hold every field of
the final state on a strided lattice of columns (all levels; the lattice includes both outermost
rows / columns), plus each field's max-norm over the whole grid, which the tests use as the scale
of the relative error.  tests/test_gpu_config_sizes.py rebuilds the same cases on the GPU.

    python tests/golden/generate_config_sizes.py [c2] [c3] [c5k]
"""
import os
import sys
import time
from datetime import datetime, timedelta

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import boundary as ob  # noqa: E402
from oracle import isentropic as oi  # noqa: E402
from tests import helpers as hp  # noqa: E402

CASES = {
    # name: (nx, ny, nz, steps, stride, topo_seconds, damp_depth)
    "c2": (161, 161, 60, 100, 8, 1800.0, 15),
    "c5k": (512, 512, 64, 3, 16, 5.0, 15),
    "c3": (256, 256, 60, 20, 16, 60.0, 15),
}


def lattice(n, stride):
    idx = list(range(0, n, stride))
    if idx[-1] != n - 1:
        idx.append(n - 1)
    return np.array(idx)


def dry_case(nx, ny, nz, topo_seconds):
    """Same construction as tests/test_gpu_config_sizes.py::dry_case (config 2's 2.2 km spacing)."""
    from tasmania_b200.grid import Grid, Topography, gaussian_profile, isentropic_state_from_brunt_vaisala

    hx, hy = 1.1 * (nx - 1), 1.1 * (ny - 1)
    x = np.linspace(-hx, hx, nx)
    y = np.linspace(-hy, hy, ny)
    topo = Topography(gaussian_profile(x, y, 500.0, 50.0, 50.0), timedelta(seconds=topo_seconds))
    grid = Grid((-hx, hx), nx, (-hy, hy), ny, (400.0, 280.0), nz, units_to_m=1e3, topography=topo)
    return grid, isentropic_state_from_brunt_vaisala(grid, 22.5, 0.0, 0.015)


def run_dry(name):
    nx, ny, nz, nsteps, stride, topo_seconds, damp_depth = CASES[name]
    S, SU, SV, U, V, MTG = oi.S, oi.SU, oi.SV, oi.U, oi.V, oi.MTG
    P, EXN, H = hp.P, hp.EXN, hp.H
    dt = timedelta(seconds=5)
    grid, np_state = dry_case(nx, ny, nz, topo_seconds)
    pt = float(np_state[P][0, 0, 0])
    ogrid = oi.Grid(nx, ny, nz, grid.dx, grid.dy, grid.dz, grid.z_on_interface_levels, grid.z)
    ohb = ob.Relaxed(nx, ny, nz, 3, 6)
    ostate = {n: v.copy() for n, v in np_state.items()}
    ostate["time"] = datetime(2000, 1, 1)
    ohb.reference_state = {n: v.copy() for n, v in np_state.items()}
    otopo = hp.Topography(grid.topography.steady_profile, topo_seconds)
    odyc = oi.IsentropicDycore(ogrid, ohb, otopo, scheme="rk3ws_si", flux="fifth_order_upwind",
                               pt=pt, eps=0.5, damp=True, damp_depth=damp_depth, damp_max=5e-4)
    for step in range(nsteps):
        otopo.update((step + 1) * dt)
        out = odyc(ostate, {}, dt)
        new = {n: out[n].copy() for n in (S, SU, U, SV, V)}
        new["time"] = out["time"]
        for n in (P, EXN, H, MTG):
            new[n] = ostate[n].copy()
        oi.refresh_diagnostics(ogrid, otopo(), new[S], pt, new[P], new[EXN], new[MTG], new[H])
        ostate = new
    return (nx, ny, nz, nsteps, stride), {n: ostate[n] for n in (S, SU, SV, U, V, MTG, P, EXN, H)}, 3


def run_moist(name):
    from oracle import moist_model as mm

    nx, ny, nz, nsteps, stride, topo_seconds, damp_depth = CASES[name]
    dt = timedelta(seconds=5)
    grid, np_state = hp.moist_case(nx, ny, nz, topo_seconds=topo_seconds,
                                   half_width_km=(1.1 * (nx - 1), 1.1 * (ny - 1)))
    ogrid = oi.Grid(nx, ny, nz, grid.dx, grid.dy, grid.dz, grid.z_on_interface_levels, grid.z)
    ohb = ob.Relaxed(nx, ny, nz, 3, 6)
    ohb.reference_state = {n: v.copy() for n, v in np_state.items()}
    otopo = hp.Topography(grid.topography.steady_profile, topo_seconds)
    pt = float(np_state[mm.P][0, 0, 0])
    omodel = mm.MoistIsentropicModel(ogrid, ohb, otopo, pt, damp_depth=damp_depth)
    ost = {n: v.copy() for n, v in np_state.items()}
    ost[mm.W] = np.zeros_like(ost[mm.S])
    ost["time"] = datetime(1992, 2, 20)
    for _ in range(nsteps):
        ost = omodel.step(ost, dt)
    names3 = (mm.S, mm.SU, mm.SV, mm.U, mm.V, mm.MTG, mm.P, mm.EXN, mm.H, mm.RHO, mm.T, mm.QV,
              mm.QC, mm.QR, mm.W, mm.VT)
    fields = {n: ost[n] for n in names3}
    fields.update({n: ost[n] for n in (mm.PREC, mm.ACCPREC)})
    return (nx, ny, nz, nsteps, stride), fields, 3


def main(which):
    for name in which:
        t0 = time.time()
        dims, fields, _ = run_moist(name) if name == "c3" else run_dry(name)
        nx, ny, nz, nsteps, stride = dims
        li, lj = lattice(nx, stride), lattice(ny, stride)
        out = {"dims": np.array(dims), "li": li, "lj": lj}
        for n, a in fields.items():
            a = np.asarray(a)
            box = a[:nx, :ny, :nz] if a.ndim == 3 and a.shape[2] > 1 else a[:nx, :ny]
            out["lat_" + n] = np.ascontiguousarray(box[np.ix_(li, lj)])
            out["max_" + n] = np.array(float(np.max(np.abs(box))))
        path = os.path.join(HERE, f"config_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {dims} in {time.time() - t0:.0f} s -> {path} ({os.path.getsize(path) / 1e6:.1f} MB)",
              flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
