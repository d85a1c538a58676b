# -*- coding: utf-8 -*-
"""CUDA-graph replay of the time step (tasmania_b200.graphs) on the GPU: the graphed runs of the
dry loop (configs[1]) and of the moist SUS model (configs[2]) are bit-identical to eager stepping,
field by field, and after one buffer-rotation period a step costs a single graph launch."""
from datetime import timedelta

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dry(nx=45, ny=37, nz=12):
    from tasmania_b200.grid import Grid, Topography, gaussian_profile, isentropic_state_from_brunt_vaisala
    from tasmania_b200.isentropic_dry import IsentropicDryRun

    x, y = np.linspace(-176.0, 176.0, nx), np.linspace(-176.0, 176.0, ny)
    topo = Topography(gaussian_profile(x, y, 500.0, 50.0, 50.0), timedelta(seconds=40))
    grid = Grid((-176.0, 176.0), nx, (-176.0, 176.0), ny, (400.0, 280.0), nz, units_to_m=1e3,
                topography=topo)
    np_state = isentropic_state_from_brunt_vaisala(grid, 22.5, 0.0, 0.015)
    return IsentropicDryRun(grid, np_state, timedelta(seconds=5), damp_depth=4)


def _moist(nx=33, ny=29, nz=14):
    from tasmania_b200.isentropic_moist import IsentropicMoistSUS
    from tests import helpers as hp

    grid, np_state = hp.moist_case(nx, ny, nz)
    return IsentropicMoistSUS(grid, np_state, timedelta(seconds=5), damp_depth=4)


@pytest.mark.parametrize("build,period,nsteps", [(_dry, 2, 15), (_moist, 12, 29)])
def test_graphed_steps_equal_eager_steps_bitwise(build, period, nsteps):
    import torch

    import tasmania_b200 as tb
    from tasmania_b200.graphs import GraphedLoop

    eager, graphed = build(), build()
    loop = GraphedLoop(graphed)
    n0 = tb.lib.launch_count()
    for _ in range(nsteps):
        eager.step()
    per_step = (tb.lib.launch_count() - n0) / nsteps
    n0 = tb.lib.launch_count()
    loop.run(nsteps)
    torch.cuda.synchronize()
    assert loop.period == period
    # once every configuration is captured, steps are replays: launches recorded by the library
    # stop growing with the number of steps, the replayed ones account for the rest
    eager_launches = tb.lib.launch_count() - n0
    assert eager_launches + loop.replayed_launches >= per_step * nsteps - 1
    assert loop.replayed_launches >= per_step * (nsteps - loop.eager_steps) * 0.9
    assert graphed.state["time"] == eager.state["time"]
    assert set(graphed.state) == set(eager.state)
    for n, v in eager.state.items():
        if n == "time":
            continue
        a, b = tb.to_numpy(graphed.state[n]), tb.to_numpy(v)
        assert np.isfinite(b).all(), n
        np.testing.assert_array_equal(a, b, err_msg=n)
    # graphed and eager steps can be interleaved: the loop keys on the buffer configuration
    eager.step()
    graphed.step()
    loop.step()
    eager.step()
    for n in ("air_isentropic_density", "x_momentum_isentropic"):
        np.testing.assert_array_equal(tb.to_numpy(graphed.state[n]), tb.to_numpy(eager.state[n]))
