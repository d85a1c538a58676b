# -*- coding: utf-8 -*-
"""NetCDFMonitor on the GPU: ``store`` snapshots the state on the device and transfers it on a side
stream while the model keeps stepping (and recycling the very arrays that were stored); the file
holds, for every record, exactly what ``to_numpy`` returned at the time of ``store``."""
import os
from datetime import timedelta

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_store_while_stepping(tmp_path):
    import tasmania_b200 as tb
    from tasmania_b200.graphs import GraphedLoop
    from tasmania_b200.iox import NetCDFMonitor, grid_shape, load_netcdf_dataset
    from tests.test_gpu_graphs import _dry

    run = _dry(61, 53, 16)
    loop = GraphedLoop(run)
    fn = os.path.join(tmp_path, "dry.nc")
    names = ("air_isentropic_density", "x_momentum_isentropic", "x_velocity_at_u_locations",
             "air_pressure_on_interface_levels")
    mon = NetCDFMonitor(fn, run.grid, store_names=names)
    want = []
    for step in range(9):
        loop.step()
        if step % 2 == 0:
            mon.store(run.state)          # returns at once; the next steps overwrite these arrays
            if step == 4:                 # (reading back here synchronises: test only)
                want.append({n: tb.to_numpy(run.state[n]) for n in names})
            else:
                want.append(None)
    want_last = {n: tb.to_numpy(run.state[n]) for n in names}
    mon.write()
    _, _, states = load_netcdf_dataset(fn)
    assert len(states) == 5
    assert [s["time"] for s in states] == [run.init_time + (2 * k + 1) * timedelta(seconds=5) for k in range(5)]
    for rec, ref in ((2, want[2]), (4, want_last)):
        for n in names:
            shape, _ = grid_shape(run.grid, n, ref[n].shape)
            np.testing.assert_array_equal(states[rec][n], ref[n][: shape[0], : shape[1], : shape[2]], err_msg=n)
    # consecutive records differ (the flow evolves) and are finite
    assert not np.array_equal(states[0]["x_momentum_isentropic"], states[4]["x_momentum_isentropic"])
    assert all(np.isfinite(v).all() for s in states for k, v in s.items() if k != "time")
