# -*- coding: utf-8 -*-
"""Model output without stalling the step loop (SURVEY.md section 8f-4, I/O half): host-side mirror of
``tasmania.NetCDFMonitor`` (src/tasmania/utils/iox.py:L47-L296, on sympl's NetCDFMonitor) and of
``load_netcdf_dataset`` (iox.py:L299-L442) for b200 storages -- same constructor arguments (a grid
in place of the sympl ``Domain``), ``store(state)`` / ``write()``.

The reference deep-copies the state on ``store`` (iox.py:L136-L141) and converts to numpy when
writing; with the state on the device that is a device-to-host transfer of every stored field,
which at PCIe speed costs more than the time step itself (5 x 0.56 GB = 50 ms against an 8.6 ms step at
1024 x 1024 x 64).  Here ``store`` only enqueues work:

  1. on the compute stream, a device-to-device snapshot of the stored fields into one of two
     staging sets (HBM speed: microseconds to ~1 ms), so the model may overwrite its arrays at once;
  2. on a dedicated copy stream, the device-to-host transfer of that snapshot into pinned host
     buffers, overlapping the following time steps; events order it after the snapshot and keep a
     staging set from being reused before its transfer has finished.

``write()`` waits for the outstanding transfers and appends the records to the file.  Files are
NetCDF-3 (64-bit offset) written with ``scipy.io.netcdf_file`` (netCDF4 / xarray, which the
reference uses, are not available offline; both read this format): one unlimited ``time``
dimension, the reference's dimension names for the axes, a ``units`` attribute per variable.
"""
from __future__ import annotations

import os
from datetime import datetime, timedelta

import numpy as np
import torch

from tasmania_b200 import storage

# units of the model variables on this path (the `units` entries of the components'
# input / output properties in the reference)
UNITS = {
    "air_isentropic_density": "kg m^-2 K^-1",
    "x_momentum_isentropic": "kg m^-1 K^-1 s^-1",
    "y_momentum_isentropic": "kg m^-1 K^-1 s^-1",
    "x_velocity_at_u_locations": "m s^-1",
    "y_velocity_at_v_locations": "m s^-1",
    "x_velocity": "m s^-1",
    "y_velocity": "m s^-1",
    "air_pressure_on_interface_levels": "Pa",
    "exner_function_on_interface_levels": "J K^-1 kg^-1",
    "height_on_interface_levels": "m",
    "montgomery_potential": "m^2 s^-2",
    "air_density": "kg m^-3",
    "air_temperature": "K",
    "mass_fraction_of_water_vapor_in_air": "g g^-1",
    "mass_fraction_of_cloud_liquid_water_in_air": "g g^-1",
    "mass_fraction_of_precipitation_water_in_air": "g g^-1",
    "tendency_of_air_potential_temperature": "K s^-1",
    "raindrop_fall_velocity": "m s^-1",
    "precipitation": "mm hr^-1",
    "accumulated_precipitation": "mm",
}
TIME_UNITS = {"seconds": 1.0, "minutes": 60.0, "hours": 3600.0, "days": 86400.0}


def grid_shape(grid, name, storage_shape):
    """Shape of ``name`` on the grid (the ``grid_shape`` of get_dataarray_3d in the reference) and
    its dimension names."""
    nx, ny, nz = grid.nx, grid.ny, grid.nz
    stg_x = "at_u_locations" in name or "at_uv_locations" in name
    stg_y = "at_v_locations" in name or "at_uv_locations" in name
    stg_z = "on_interface_levels" in name
    nk = 1 if storage_shape[2] == 1 else (nz + 1 if stg_z else nz)
    shape = (nx + 1 if stg_x else nx, ny + 1 if stg_y else ny, nk)
    dims = ("x_at_u_locations" if stg_x else "x", "y_at_v_locations" if stg_y else "y",
            "surface" if nk == 1 and nz != 1 else
            ("air_potential_temperature_on_interface_levels" if stg_z else "air_potential_temperature"))
    return shape, dims


def _flat(arr):
    t = arr.t
    return t._base if t._base is not None else t


class NetCDFMonitor:
    def __init__(self, filename, grid, time_units="seconds", store_names=None, write_on_store=False,
                 aliases=None):
        if time_units not in TIME_UNITS:
            raise ValueError(f"time_units must be one of {sorted(TIME_UNITS)}")
        self.filename, self.grid, self.time_units = filename, grid, time_units
        self._store_names = tuple(store_names) if store_names is not None else None
        self._write_on_store = bool(write_on_store)
        self._aliases = dict(aliases or {})
        self._pending = []   # (time, {name: (pinned host tensor, like, grid shape, dims)}, event)
        self._host_pool = {}  # (name, shape, pinned) -> host buffers free for reuse
        self._staging = [None, None]   # device staging sets: {name: flat tensor}
        self._staging_free = [None, None]  # event: the D2H copy out of the set has finished
        self._nstored = 0
        self._copy_stream = None
        self._epoch = None
        self._created = False
        self.records_written = 0

    # ---- store: enqueue snapshot + transfer, return immediately
    def _names(self, state):
        names = [n for n in state if n != "time"]
        if self._store_names is not None:
            missing = [n for n in self._store_names if n not in state]
            if missing:
                raise KeyError(f"state has no {missing}")
            names = [n for n in names if n in self._store_names]
        return names

    def store(self, state):
        if "time" not in state:
            raise KeyError("the state must carry a 'time' entry")
        names = self._names(state)
        on_device = any(state[n].t.is_cuda for n in names)
        b = self._nstored % 2
        self._nstored += 1
        host, meta = {}, {}
        for n in names:
            flat = _flat(state[n])
            assert flat.is_contiguous(), "host copies mirror the padded allocation: it must be contiguous"
            # pinned buffers are recycled once write() has consumed them: allocating pinned memory
            # synchronises with the device, which would stall the step loop on every store()
            pool = self._host_pool.setdefault((n, tuple(flat.shape), bool(on_device)), [])
            host[n] = pool.pop() if pool else torch.empty(flat.shape, dtype=flat.dtype, device="cpu",
                                                          pin_memory=bool(on_device))
            t = state[n].t
            view = (tuple(t.shape), tuple(t.stride()), t.storage_offset()) if t._base is not None else None
            meta[n] = (view, ) + grid_shape(self.grid, n, state[n].shape)
        event = None
        if on_device:
            main = torch.cuda.current_stream()
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream()
            if self._staging_free[b] is not None:
                main.wait_event(self._staging_free[b])  # its last transfer must be over
            stage = self._staging[b]
            if stage is None or any(n not in stage or stage[n].shape != _flat(state[n]).shape for n in names):
                stage = self._staging[b] = {n: torch.empty_like(_flat(state[n])) for n in names}
            for n in names:
                stage[n].copy_(_flat(state[n]), non_blocking=True)  # D2D snapshot, compute stream
            snap = torch.cuda.Event()
            snap.record(main)
            self._copy_stream.wait_event(snap)
            with torch.cuda.stream(self._copy_stream):
                for n in names:
                    host[n].copy_(stage[n], non_blocking=True)
                event = torch.cuda.Event()
                event.record()
            self._staging_free[b] = event
        else:  # host storages (tests of the host logic): plain copies
            for n in names:
                host[n].copy_(_flat(state[n]))
        self._pending.append((state["time"], host, meta, event))
        if self._write_on_store:
            self.write()

    # ---- write: drain the transfers, append the records
    @staticmethod
    def _logical(host_flat, view, shape):
        """The (i, j, k) array of a field from the host copy of its padded allocation."""
        arr = host_flat if view is None else torch.as_strided(host_flat, *view)
        return arr[: shape[0], : shape[1], : shape[2]].numpy()

    def write(self):
        import scipy.io

        if not self._pending:
            return
        for _, _, _, event in self._pending:
            if event is not None:
                event.synchronize()
        g = self.grid
        exists = self._created and os.path.exists(self.filename)
        nc = scipy.io.netcdf_file(self.filename, "a" if exists else "w", version=2)
        try:
            if not exists:
                self._epoch = self._pending[0][0]
                nc.createDimension("time", None)
                tv = nc.createVariable("time", "f8", ("time",))
                tv.units = f"{self.time_units} since {self._epoch:%Y-%m-%d %H:%M:%S}"
                coords = {
                    "x": g.x, "x_at_u_locations": np.concatenate((g.x - 0.5 * g.dx, [g.x[-1] + 0.5 * g.dx])),
                    "y": g.y, "y_at_v_locations": np.concatenate((g.y - 0.5 * g.dy, [g.y[-1] + 0.5 * g.dy])),
                    "air_potential_temperature": g.z,
                    "air_potential_temperature_on_interface_levels": g.z_on_interface_levels,
                    "surface": np.array([g.z_on_interface_levels[-1]]),
                }
                for dim, vals in coords.items():
                    nc.createDimension(dim, len(vals))
                    cv = nc.createVariable(dim, "f8", (dim,))
                    cv[:] = np.asarray(vals, dtype=float)
                    cv.units = "m" if dim[0] in "xy" else "K"
                nc.grid_type = "numerical"
            rec = nc.variables["time"].shape[0]
            for time, host, meta, _ in self._pending:
                nc.variables["time"][rec] = (time - self._epoch).total_seconds() / TIME_UNITS[self.time_units]
                for n, flat in host.items():
                    view, shape, dims = meta[n]
                    out_name = self._aliases.get(n, n)
                    if out_name not in nc.variables:
                        var = nc.createVariable(out_name, "f8", ("time",) + dims)
                        var.units = UNITS.get(n, "1")
                    nc.variables[out_name][rec] = self._logical(flat, view, shape)
                rec += 1
                self.records_written += 1
        finally:
            nc.close()
        self._created = True
        for _, host, _, event in self._pending:  # written: the host buffers go back to the pool
            on_device = event is not None
            for n, flat in host.items():
                self._host_pool.setdefault((n, tuple(flat.shape), on_device), []).append(flat)
        self._pending = []


def load_netcdf_dataset(filename):
    """-> (coordinates dict, grid_type, list of states): every record as a dict name -> numpy array
    (+ ``"time"``), the counterpart of iox.py:L299-L442 at the raw-array level."""
    import scipy.io

    nc = scipy.io.netcdf_file(filename, "r", mmap=False)
    try:
        tvar = nc.variables["time"]
        units = tvar.units.decode() if isinstance(tvar.units, bytes) else tvar.units
        unit, _, epoch = units.partition(" since ")
        t0 = datetime.strptime(epoch, "%Y-%m-%d %H:%M:%S")
        times = [t0 + timedelta(seconds=float(v) * TIME_UNITS[unit]) for v in tvar[:]]
        coords = {n: v[:].copy() for n, v in nc.variables.items() if v.dimensions == (n,) and n != "time"}
        gt = nc.grid_type.decode() if isinstance(nc.grid_type, bytes) else nc.grid_type
        states = []
        for r, time in enumerate(times):
            st = {"time": time}
            for n, v in nc.variables.items():
                if v.dimensions and v.dimensions[0] == "time" and n != "time":
                    st[n] = np.array(v[r])
            states.append(st)
    finally:
        nc.close()
    return coords, gt, states


def to_device_state(state, grid, device=None):
    """A loaded record back on b200 storages of the model's storage shape (nx + 1, ny + 1, nz + 1)."""
    out = {}
    for n, v in state.items():
        if n == "time":
            out[n] = v
            continue
        full = np.zeros((grid.nx + 1, grid.ny + 1, 1 if v.shape[2] == 1 and grid.nz != 1 else grid.nz + 1))
        full[: v.shape[0], : v.shape[1], : v.shape[2]] = v
        out[n] = storage.as_storage(full, device=device)
    return out
