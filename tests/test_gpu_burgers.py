# -*- coding: utf-8 -*-
"""BASELINE config 1 on the GPU: 2-D Burgers, Zhao initial condition, 101x101, third-order
advection, RK3WS, Dirichlet (analytic) boundaries, 100 steps -- against the oracle's restatement
of BurgersDynamicalCore.  The stencil has no transcendental call and the rim values are host
numpy on both sides, so the comparison is bit-exact."""
from datetime import datetime, timedelta

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scheme,flux,nb,nsteps", [("rk3ws", "third_order", 2, 100),
                                                   ("rk2", "fifth_order", 3, 20),
                                                   ("forward_euler", "first_order", 1, 20)])
def test_burgers_dycore_vs_oracle(scheme, flux, nb, nsteps):
    import tasmania_b200 as tb
    from oracle import burgers as ob
    from tasmania_b200.boundary import Dirichlet
    from tasmania_b200.burgers import BurgersDynamicalCore, ZhaoSolutionFactory
    from tasmania_b200.grid import Grid

    nx = ny = 101
    eps = 0.01
    grid = Grid((0.0, 1.0), nx, (0.0, 1.0), ny, (0.0, 1.0), 1)
    t0 = datetime(2000, 1, 1)
    zsf = ZhaoSolutionFactory(t0, eps)
    dt = timedelta(seconds=0.001)

    host = {n: np.zeros((nx, ny, 1)) for n in ("x_velocity", "y_velocity")}
    for n in host:
        host[n][...] = zsf(t0, grid, field_name=n)
        # the mirror's analytic solution is the oracle's (state.py:L97-L152)
        np.testing.assert_array_equal(host[n], ob.zhao_solution(0.0, grid.x, grid.y, eps, n))
    hb = Dirichlet(nx, ny, 1, nb, core=zsf, grid=grid)
    state = {n: tb.as_storage(a) for n, a in host.items()}
    state["time"] = t0
    hb.reference_state = state
    dyc = BurgersDynamicalCore(grid, hb, scheme, flux)

    odyc = ob.BurgersDycore(
        nx, ny, grid.dx, grid.dy, nb, scheme=scheme, flux=flux,
        dirichlet=lambda time, sx, sy, name: ob.zhao_solution(
            (time - t0).total_seconds(), grid.x[sx], grid.y[sy], eps, name))
    ostate = {n: a.copy() for n, a in host.items()}
    ostate["time"] = t0

    for _ in range(nsteps):
        out = dyc(state, {}, dt)
        state = {"x_velocity": out["x_velocity"].copy(), "y_velocity": out["y_velocity"].copy(),
                 "time": out["time"]}
        oout = odyc(ostate, {}, dt)
        ostate = {"x_velocity": oout["x_velocity"].copy(), "y_velocity": oout["y_velocity"].copy(),
                  "time": oout["time"]}
    assert state["time"] == ostate["time"]
    for n in ("x_velocity", "y_velocity"):
        np.testing.assert_array_equal(tb.to_numpy(state[n]), ostate[n], err_msg=n)
    # sanity: the advective core alone (the reference never couples the diffusion component
    # into the dycore) stays close to the viscous analytic solution
    exact = zsf(state["time"], grid, field_name="x_velocity")
    assert np.abs(tb.to_numpy(state["x_velocity"]) - exact).max() < 0.25 * np.abs(exact).max()
