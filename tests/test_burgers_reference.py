# -*- coding: utf-8 -*-
"""BASELINE configs[0] (2-D Burgers, Zhao initial condition, 101x101, third-order advection, RK3WS,
Dirichlet boundaries from the analytic solution) pinned on the REFERENCE run in place (skipped
where /root/reference is absent): the reference's own numpy backend --
``BurgersDynamicalCore.stage_array_call`` (src/tasmania/burgers/dynamics/dycore.py:L158-L173, unbound
on a stand-in object holding the reference's real RK3WS ``BurgersStepper`` and ``Dirichlet`` boundary
with the real ``ZhaoSolutionFactory`` core), chained as framework/dycore.py:L455-L458 does -- against
the oracle's ``BurgersDycore`` and the b200 mirror's host-side pieces, bit for bit over 100 steps.
"""
import types
from datetime import datetime, timedelta

import numpy as np
import pytest

from oracle import burgers as ob
from tests.golden import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference tree not mounted")

NAMES = ("x_velocity", "y_velocity")


def reference_run(nx, ny, nb, scheme, flux, eps, dt, nsteps, t0):
    refload.install_framework()
    DataArray = refload.DataArray
    dom = refload.load("tasmania.domain.domain")
    refload.load("tasmania.domain.subclasses.horizontal_boundaries.dirichlet")
    refload.load("tasmania.domain.subclasses.topographies.flat")
    bstate = refload.load("tasmania.burgers.state")
    bst = refload.load("tasmania.burgers.dynamics.stepper")
    bdy = refload.load("tasmania.burgers.dynamics.dycore")
    for m in ("forward_euler", "rk2", "rk3ws"):
        refload.load("tasmania.burgers.dynamics.subclasses.stepper." + m)
    for m in ("first_order", "second_order", "third_order", "fourth_order", "fifth_order", "sixth_order"):
        refload.load("tasmania.burgers.dynamics.subclasses.advection." + m)
    opts = refload.load("tasmania.framework.options")
    core = bstate.ZhaoSolutionFactory(t0, DataArray(eps, attrs={"units": "m^2 s^-1"}))
    domain = dom.Domain(
        DataArray([0, 1], dims="x", attrs={"units": "m"}), nx,
        DataArray([0, 1], dims="y", attrs={"units": "m"}), ny,
        DataArray([0, 1], dims="z", attrs={"units": "1"}), 1,
        horizontal_boundary_type="dirichlet", nb=nb, horizontal_boundary_kwargs={"core": core},
        backend="numpy")
    grid = domain.numerical_grid
    stepper = bst.BurgersStepper.factory(scheme, grid.grid_xy, nb, flux, backend="numpy",
                                         backend_options=opts.BackendOptions(),
                                         storage_options=opts.StorageOptions())
    me = types.SimpleNamespace(_stepper=stepper, horizontal_boundary=domain.horizontal_boundary)
    state = {n: np.array(core(t0, grid, field_name=n)) for n in NAMES}
    initial = {n: v.copy() for n, v in state.items()}
    state["time"] = t0
    # enforce_raw only touches fields that have a reference value (horizontal_boundary.py:L322-L331)
    domain.horizontal_boundary.reference_state = {
        **{n: DataArray(v.copy(), attrs={"units": "m s^-1"}) for n, v in initial.items()}, "time": t0}
    outs = [{n: np.zeros((nx, ny, 1)) for n in NAMES} for _ in range(stepper.stages)]
    for _ in range(nsteps):
        cur = state
        for stage in range(stepper.stages):
            bdy.BurgersDynamicalCore.stage_array_call(me, stage, cur, {}, dt, outs[stage])
            cur = outs[stage]
        state = {n: cur[n].copy() for n in NAMES}
        state["time"] = cur["time"]
    x = np.asarray(grid.x.to_units("m").values)
    y = np.asarray(grid.y.to_units("m").values)
    dx, dy = grid.dx.to_units("m").values.item(), grid.dy.to_units("m").values.item()
    return initial, state, (x, y, dx, dy), core, grid


@pytest.mark.parametrize("scheme,flux,nb,nsteps", [("rk3ws", "third_order", 2, 100),
                                                   ("rk2", "fifth_order", 3, 15),
                                                   ("forward_euler", "first_order", 1, 15),
                                                   # the even orders, so that all six advection
                                                   # schemes are pinned by execution
                                                   ("rk3ws", "second_order", 1, 10),
                                                   ("rk2", "fourth_order", 2, 10),
                                                   ("forward_euler", "sixth_order", 3, 10)])
def test_reference_burgers_dycore_equals_oracle(scheme, flux, nb, nsteps):
    nx = ny = 101
    eps, t0, dt = 0.01, datetime(2000, 1, 1), timedelta(seconds=0.001)
    initial, final, (x, y, dx, dy), _, _ = reference_run(nx, ny, nb, scheme, flux, eps, dt, nsteps, t0)
    for n in NAMES:  # the Zhao state, src/tasmania/burgers/state.py:L97-L152
        np.testing.assert_array_equal(initial[n], ob.zhao_solution(0.0, x, y, eps, n), err_msg=n)
    odyc = ob.BurgersDycore(
        nx, ny, dx, dy, nb, scheme=scheme, flux=flux,
        dirichlet=lambda time, sx, sy, name: ob.zhao_solution((time - t0).total_seconds(), x[sx], y[sy],
                                                               eps, name))
    ostate = {n: v.copy() for n, v in initial.items()}
    ostate["time"] = t0
    for _ in range(nsteps):
        out = odyc(ostate, {}, dt)
        ostate = {n: out[n].copy() for n in NAMES}
        ostate["time"] = out["time"]
    assert abs((final["time"] - ostate["time"]).total_seconds()) < 1e-3 * nsteps  # stage-wise time labels
    for n in NAMES:
        np.testing.assert_array_equal(final[n], ostate[n], err_msg=n)
    assert float(np.abs(final["x_velocity"] - initial["x_velocity"]).max()) > 0.0


def test_mirror_host_pieces_equal_reference():
    """tasmania_b200's host-side Burgers set-up -- grid spacing, the analytic solution factory, the
    rim slabs the Dirichlet boundary evaluates -- against the reference's objects."""
    from tasmania_b200.burgers import ZhaoSolutionFactory
    from tasmania_b200.grid import Grid

    nx, ny, nb, eps, t0 = 101, 81, 2, 0.01, datetime(2000, 1, 1)
    _, _, (x, y, dx, dy), core, rgrid = reference_run(nx, ny, nb, "rk3ws", "third_order", eps,
                                                      timedelta(seconds=0.001), 0, t0)
    grid = Grid((0.0, 1.0), nx, (0.0, 1.0), ny, (0.0, 1.0), 1)
    assert (grid.dx, grid.dy) == (dx, dy)
    np.testing.assert_array_equal(grid.x, x)
    zsf = ZhaoSolutionFactory(t0, eps)
    later = t0 + timedelta(seconds=0.37)
    for n in NAMES:
        np.testing.assert_array_equal(zsf(later, grid, field_name=n), np.array(core(later, rgrid, field_name=n)))
        for sx, sy in ((slice(0, nb), slice(0, ny)), (slice(nx - nb, nx), slice(0, ny)),
                       (slice(nb, nx - nb), slice(0, nb)), (slice(nb, nx - nb), slice(ny - nb, ny))):
            np.testing.assert_array_equal(zsf(later, grid, sx, sy, n), np.array(core(later, rgrid, sx, sy, n)))
