# -*- coding: utf-8 -*-
"""The UNMODIFIED reference dry dynamical core, time-stepped on two backends side by side:
``backend="numpy"`` (the reference's own CPU path) and ``backend="b200"`` through the plugin --
per-stencil kernels, or (default) the fused stage behind the reference's own
``IsentropicDynamicalCore.stage_array_call_dry`` (plugin.install(fused_stage=True)).

Everything numerical is the reference's: Domain, Relaxed boundary, state builder, RK3WSSI
prognostic, Rayleigh damper, HorizontalVelocity, IsentropicDiagnostics, stage_array_call_dry, the
stage chaining of framework/dycore.py:L455-L458 (the sympl wrappers around them cannot be
instantiated offline: tests/golden/refload.py).  north_star's bound: 1e-12 relative on every
field after the steps (bit-exact when the oracle-backed stub stands in for the library).

    python tests/ref_dycore_steps.py [--steps 10] [--stub] [--per-stencil] [--nx 41 --ny 37 --nz 12]

``--stub``: the oracle-backed C-ABI stub instead of libtasmania_b200.so (build container, no GPU).
Reference location: $TASMANIA_REFERENCE (default /root/reference; on the GPU box the copy staged by
baseline/stage_reference.sh under baseline/_ref).
"""
import argparse
import contextlib
import os
import sys
import types
from datetime import datetime, timedelta

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import refload  # noqa: E402

refload.install_framework()

import generate_golden as gg  # noqa: E402

S, SU, SV = gg.S, gg.SU, gg.SV
U, V, MTG = "x_velocity_at_u_locations", "y_velocity_at_v_locations", "montgomery_potential"
P, EXN, H = gg.P, "exner_function_on_interface_levels", "height_on_interface_levels"
NB, NR = 3, 6
DT = timedelta(seconds=5)
OUTNAMES = (S, SU, U, SV, V)


QNAMES = ("mass_fraction_of_water_vapor_in_air", "mass_fraction_of_cloud_liquid_water_in_air",
          "mass_fraction_of_precipitation_water_in_air")


def run(backend, nsteps, nx, ny, nz, clock=None, damp_depth=4, topo_seconds=20, moist=False, boundary="relaxed"):
    """``nsteps`` RK3WS + fifth-order-upwind steps of the mountain-flow case on ``backend``;
    returns the final state as numpy arrays.  ``clock`` (a dict with a "budget" in seconds): the
    step loop is timed (set-up excluded) and stops once the budget is spent -- how ``bench.py
    --impl reference`` times the reference's own numpy backend; "steps" and "seconds" are set."""
    from tasmania.framework import allocators as ta
    from tasmania.framework.generic_functions import to_numpy

    DataArray, da = refload.DataArray, gg.da
    dom = refload.load("tasmania.domain.domain")
    refload.load("tasmania.domain.subclasses.horizontal_boundaries.relaxed")
    refload.load("tasmania.domain.subclasses.horizontal_boundaries.periodic")
    refload.load("tasmania.domain.subclasses.topographies.gaussian")
    domain = dom.Domain(
        DataArray([-176, 176], dims="x", attrs={"units": "km"}), nx,
        DataArray([-176, 176], dims="y", attrs={"units": "km"}), ny,
        DataArray([400, 280], dims="z", attrs={"units": "K"}), nz,
        horizontal_boundary_type=boundary, nb=NB,
        horizontal_boundary_kwargs={"nr": NR} if boundary == "relaxed" else {},
        backend=backend, topography_type="gaussian",
        topography_kwargs={"time": timedelta(seconds=topo_seconds), "max_height": da(0.5, "km"),
                           "width_x": da(50.0, "km"), "width_y": da(50.0, "km"), "smooth": False})
    grid = domain.numerical_grid
    # ``boundary="periodic"``: (nx, ny) is the physical grid, the fields live on the numerical one
    # (NB ghost points a side, periodic.py:L44-L50)
    nx, ny = grid.nx, grid.ny
    shape = (nx + 1, ny + 1, nz + 1)
    st = refload.load("tasmania.isentropic.state")
    state = st.get_isentropic_state_from_brunt_vaisala_frequency(
        grid, datetime(2000, 1, 1), da(22.5, "m s^-1"), da(0.0, "m s^-1"), da(0.015, "s^-1"),
        moist=moist, relative_humidity=0.95, backend=backend, storage_shape=shape)
    if moist:  # seed cloud water and rain so that all three constituents are advected (same bits on every backend)
        i, j, k = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
        blob = np.exp(-((i - 0.4 * nx) ** 2 + (j - 0.5 * ny) ** 2) / (0.02 * nx * ny) - ((k - 0.7 * nz) / (0.2 * nz)) ** 2)
        blob[nx:, :, :] = blob[:, ny:, :] = blob[:, :, nz:] = 0.0
        for n, amp in ((QNAMES[1], 8e-4), (QNAMES[2], 3e-4)):
            state[n] = DataArray(ta.as_storage(backend, data=amp * blob), attrs={"units": "g g^-1"})
    hb = domain.horizontal_boundary
    hb.reference_state = state
    dyc = refload.load("tasmania.isentropic.dynamics.dycore")
    refload.load("tasmania.isentropic.dynamics.subclasses.prognostics.utils")
    refload.load("tasmania.isentropic.dynamics.subclasses.prognostics.rk3ws_si")
    refload.load("tasmania.isentropic.dynamics.subclasses.minimal_horizontal_fluxes.fifth_order_upwind")
    refload.load("tasmania.dwarfs.subclasses.vertical_dampers.rayleigh")
    prog = refload.load("tasmania.isentropic.dynamics.prognostic")
    diagm = refload.load("tasmania.isentropic.dynamics.diagnostics")
    vd = refload.load("tasmania.dwarfs.vertical_damping")
    dd = refload.load("tasmania.dwarfs.diagnostics")
    opts = refload.load("tasmania.framework.options")
    bo, so = opts.BackendOptions, opts.StorageOptions
    pt = float(to_numpy(state[P].data)[0, 0, 0])
    prognostic = prog.IsentropicPrognostic.factory(
        "rk3ws_si", "fifth_order_upwind", domain, moist, backend=backend, backend_options=bo(),
        storage_shape=shape, storage_options=so(), pt=da(pt, "Pa"), eps=0.5)
    damper = vd.VerticalDamping.factory("rayleigh", grid, damp_depth, 5e-4, backend=backend, backend_options=bo(),
                                        storage_shape=shape, storage_options=so())
    velocity = dd.HorizontalVelocity(grid, staggering=True, backend=backend, backend_options=bo(),
                                     storage_options=so())
    diagnostics = diagm.IsentropicDiagnostics(grid, backend=backend, backend_options=bo(),
                                              storage_shape=shape, storage_options=so())

    def zeros():
        return ta.zeros(backend, shape=shape)

    outnames = OUTNAMES + (QNAMES if moist else ())
    water = {}
    if moist:
        water["_water_constituent"] = dd.WaterConstituent(
            grid, clipping=True, backend=backend, backend_options=bo(), storage_options=so())
        water.update({f"_{q}_{t}": zeros() for q in ("sqv", "sqc", "sqr") for t in ("now", "int", "new")})
    me = types.SimpleNamespace(  # the attributes stage_array_call_dry / _moist read from the dycore object
        horizontal_boundary=hb, backend=backend, grid=grid, storage_options=so(), _moist=moist, **water,
        fast_tendency_component=None, fast_diagnostic_component=None,
        output_properties={k: {"units": state[k].attrs["units"]} for k in outnames},
        _damp=True, _damp_at_every_stage=True, stages=prognostic.stages, _prognostic=prognostic,
        _damper=damper, _velocity_components=velocity, _s_ref=zeros(), _su_ref=zeros(),
        _sv_ref=zeros(), _s_now=None, _su_now=None, _sv_now=None)
    cur = {k: state[k].data for k in (S, MTG, SU, U, SV, V, P, EXN, H) + (QNAMES if moist else ())}
    cur["time"] = state["time"]
    stage_outs = [{k: zeros() for k in outnames} for _ in range(prognostic.stages - 1)]
    spare = {k: zeros() for k in outnames}
    stage_call = (dyc.IsentropicDynamicalCore.stage_array_call_moist if moist
                  else dyc.IsentropicDynamicalCore.stage_array_call_dry)
    import time as _time

    t_start = _time.perf_counter()
    for step in range(nsteps):
        grid.update_topography((step + 1) * DT)
        outs = stage_outs + [spare]
        st_in = cur
        for stage in range(prognostic.stages):  # stage chaining of framework/dycore.py:L455-L458
            stage_call(me, stage, st_in, {}, DT, outs[stage])
            st_in = dict(outs[stage])
            st_in.setdefault(MTG, cur[MTG])
        new = {k: spare[k] for k in outnames}
        new["time"] = cur["time"] + DT
        for k in (P, EXN, H, MTG):
            new[k] = cur[k]
        spare = {k: cur[k] for k in outnames}
        # the role dv plays in driver_namelist_sus.py:L188-L199
        diagnostics.get_diagnostic_variables(new[S], pt, new[P], new[EXN], new[MTG], new[H])
        cur = new
        if clock is not None:
            clock["steps"], clock["seconds"] = step + 1, _time.perf_counter() - t_start
            if clock["seconds"] > clock["budget"]:
                break
    return {k: np.array(to_numpy(v)) for k, v in cur.items() if k != "time"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--stub", action="store_true")
    ap.add_argument("--per-stencil", action="store_true")
    ap.add_argument("--moist", action="store_true", help="the moist dycore stage (stage_array_call_moist)")
    ap.add_argument("--periodic", action="store_true", help="the reference's Periodic boundary instead of Relaxed")
    ap.add_argument("--nx", type=int, default=41)
    ap.add_argument("--ny", type=int, default=37)
    ap.add_argument("--nz", type=int, default=12)
    args = ap.parse_args()

    import tasmania_b200 as tb
    from tasmania_b200 import plugin

    report = plugin.install(fused_stage=not args.per_stencil)
    assert args.per_stencil or "IsentropicDynamicalCore.stage_array_call_dry" in report.get("fused", []), report
    if args.moist and not args.per_stencil:
        assert "IsentropicDynamicalCore.stage_array_call_moist" in report.get("fused", []), report
    boundary = "periodic" if args.periodic else "relaxed"
    want = run("numpy", args.steps, args.nx, args.ny, args.nz, moist=args.moist, boundary=boundary)
    if args.stub:
        from tests.abi_oracle import OracleStub
        from tests.abi_stub import stubbed_library

        ctx = stubbed_library(OracleStub)
    else:
        ctx = contextlib.nullcontext()
    with ctx as stub:
        n0 = None if args.stub else tb.lib.launch_count()
        got = run("b200", args.steps, args.nx, args.ny, args.nz, moist=args.moist, boundary=boundary)
        if args.stub:
            fused_calls = stub.count("tb200_isentropic_stage_moist" if args.moist else "tb200_isentropic_stage_dry")
        else:
            fused_calls = None
            # (the reference's Periodic class wraps by slice assignment: storage copies, not library launches)
            floor = 10 if (not args.per_stencil or args.periodic) else 30
            assert tb.lib.launch_count() - n0 >= floor * args.steps, (tb.lib.launch_count() - n0, floor)
    if args.stub and not args.per_stencil:
        assert fused_calls == 3 * args.steps, fused_calls
    worst = {}
    nx, ny, nz = (n - 1 for n in want[S].shape)  # the numerical grid
    for k in (S, SU, SV, U, V, MTG, P, EXN, H) + (QNAMES if args.moist else ()):
        a, b = got[k][: nx + 1, : ny + 1, : nz + 1], want[k][: nx + 1, : ny + 1, : nz + 1]
        scale = float(np.max(np.abs(b)))
        worst[k] = float(np.max(np.abs(a - b))) / scale if scale > 0 else float(np.max(np.abs(a - b)))
    print("relative errors:", {k: float(f"{v:.2e}") for k, v in worst.items()})
    tol = 0.0 if args.stub else 1e-12
    assert max(worst.values()) <= tol, worst
    assert float(np.max(np.abs(want[SV]))) > 1e-6  # the flow developed
    if args.moist:
        assert all(float(np.max(np.abs(want[q]))) > 0.0 for q in QNAMES)
    print("REF-DYCORE-STEPS-OK", ("per-stencil" if args.per_stencil else "fused") + ("-moist" if args.moist else "")
          + ("-periodic" if args.periodic else ""),
          args.steps)


if __name__ == "__main__":
    main()
