# -*- coding: utf-8 -*-
"""tasmania_b200.iox on host storages: what ``store`` + ``write`` put in the file is what the
state held at the time of ``store`` (even if the model overwrote its arrays afterwards), records
append across ``write`` calls, and ``load_netcdf_dataset`` gives the states back."""
import os
from datetime import datetime, timedelta

import numpy as np
import pytest

from tests import helpers as hp
from tests.abi_stub import stubbed_library


def test_store_write_load_roundtrip(tmp_path):
    import tasmania_b200 as tb
    from tasmania_b200.iox import NetCDFMonitor, grid_shape, load_netcdf_dataset, to_device_state

    nx, ny, nz = 17, 15, 8
    grid, np_state = hp.moist_case(nx, ny, nz)
    fn = os.path.join(tmp_path, "out.nc")
    with stubbed_library():
        state = {n: tb.as_storage(v) for n, v in np_state.items()}
        state["time"] = datetime(1992, 2, 20)
        mon = NetCDFMonitor(fn, grid, time_units="minutes",
                            aliases={"air_isentropic_density": "s"})
        mon.store(state)
        first = {n: tb.to_numpy(v) for n, v in state.items() if n != "time"}
        # the model moves on and overwrites its arrays before anything is written
        for n in first:
            state[n].t.mul_(2.0)
        state["time"] = datetime(1992, 2, 20) + timedelta(seconds=90)
        mon.store(state)
        mon.write()
        assert mon.records_written == 2
        state["time"] += timedelta(seconds=90)
        mon.store(state)
        mon.write()  # appends
        coords, grid_type, states = load_netcdf_dataset(fn)
        assert grid_type == "numerical" and len(states) == 3
        assert [s["time"] for s in states] == [datetime(1992, 2, 20) + k * timedelta(seconds=90) for k in range(3)]
        np.testing.assert_array_equal(coords["x"], grid.x)
        np.testing.assert_array_equal(coords["air_potential_temperature_on_interface_levels"],
                                      grid.z_on_interface_levels)
        for n, v in first.items():
            shape, _ = grid_shape(grid, n, v.shape)
            key = "s" if n == "air_isentropic_density" else n
            np.testing.assert_array_equal(states[0][key], v[: shape[0], : shape[1], : shape[2]], err_msg=n)
            np.testing.assert_array_equal(states[1][key], 2.0 * v[: shape[0], : shape[1], : shape[2]])
            np.testing.assert_array_equal(states[2][key], states[1][key])
        assert states[0]["x_velocity_at_u_locations"].shape == (nx + 1, ny, nz)
        assert states[0]["air_pressure_on_interface_levels"].shape == (nx, ny, nz + 1)
        assert states[0]["precipitation"].shape == (nx, ny, 1)
        # restart: a loaded record goes back onto model storages
        states[1]["air_isentropic_density"] = states[1].pop("s")
        back = to_device_state(states[1], grid)
        for n, v in first.items():
            shape, _ = grid_shape(grid, n, v.shape)
            np.testing.assert_array_equal(tb.to_numpy(back[n])[: shape[0], : shape[1], : shape[2]],
                                          2.0 * v[: shape[0], : shape[1], : shape[2]])


def test_store_names_and_errors(tmp_path):
    import tasmania_b200 as tb
    from tasmania_b200.iox import NetCDFMonitor, load_netcdf_dataset

    grid, np_state = hp.moist_case(17, 15, 8)
    fn = os.path.join(tmp_path, "sel.nc")
    with stubbed_library():
        state = {n: tb.as_storage(v) for n, v in np_state.items()}
        with pytest.raises(KeyError):
            NetCDFMonitor(fn, grid).store(state)  # no time
        state["time"] = datetime(2000, 1, 1)
        with pytest.raises(KeyError):
            NetCDFMonitor(fn, grid, store_names=("nonexistent",)).store(state)
        with pytest.raises(ValueError):
            NetCDFMonitor(fn, grid, time_units="fortnights")
        mon = NetCDFMonitor(fn, grid, store_names=("air_isentropic_density", "montgomery_potential"),
                            write_on_store=True)
        mon.store(state)
        _, _, states = load_netcdf_dataset(fn)
        assert set(states[0]) == {"time", "air_isentropic_density", "montgomery_potential"}
