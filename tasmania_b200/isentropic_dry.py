# -*- coding: utf-8 -*-
"""The dry isentropic time loop of BASELINE.json configs[1] / configs[4] on the b200 backend, as
SURVEY.md section 8d defines it (the reference snapshot ships no dry driver):

    dycore.update_topography((i + 1) * dt)
    dycore(state, {}, dt, out_state=state_new)
    IsentropicDiagnostics: p, exn, mtg, h of state_new   (the role ``dv`` plays in
                                                          driver_namelist_sus.py:L188-L199)

split into prepare / compute / finish so that tasmania_b200.graphs.GraphedLoop can replay the
compute part as a CUDA graph on launch-bound grids.
"""
from __future__ import annotations

from datetime import datetime

from tasmania_b200 import storage
from tasmania_b200.boundary import Periodic, Relaxed
from tasmania_b200.isentropic import (MTG, S, SU, SV, U, V, IsentropicDiagnostics,
                                      IsentropicDynamicalCore)

P, EXN, H = ("air_pressure_on_interface_levels", "exner_function_on_interface_levels",
             "height_on_interface_levels")


class IsentropicDryRun:
    out_names = (S, SU, U, SV, V)

    def __init__(self, grid, state, timestep, *, nb=3, nr=6, horizontal_flux_scheme="fifth_order_upwind",
                 time_integration_scheme="rk3ws_si", eps=0.5, damp=True, damp_depth=15, damp_max=5e-4,
                 damp_at_every_stage=True, init_time=None, device=None, boundary="relaxed"):
        nx, ny, nz = grid.nx, grid.ny, grid.nz
        self.grid, self.dt = grid, timestep
        self.nx, self.ny, self.nz = nx, ny, nz
        self.pt = float(state[P][0, 0, 0])
        if boundary == "relaxed":
            self.hb = Relaxed(nx, ny, nz, nb, nr=nr)
        elif boundary == "periodic":
            # ``grid`` is the numerical grid (physical + nb ghost points a side, periodic.py:L44-L50)
            # and ``state`` lives on it with wrapped ghost layers (``Periodic.get_numerical_field``)
            self.hb = Periodic(nx - 2 * nb, ny - 2 * nb, nz, nb)
        else:
            raise ValueError(f"boundary must be 'relaxed' or 'periodic', got {boundary!r}")
        self.state = {n: storage.as_storage(v, device=device) for n, v in state.items()}
        self.init_time = init_time or datetime(2000, 1, 1)
        self.state["time"] = self.init_time
        self.hb.reference_state = {n: storage.as_storage(v, device=device) for n, v in state.items()
                                   if n in (S, SU, SV, U, V)}
        self.dyc = IsentropicDynamicalCore(
            grid, self.hb, time_integration_scheme=time_integration_scheme,
            horizontal_flux_scheme=horizontal_flux_scheme,
            time_integration_properties={"pt": self.pt, "eps": eps}, damp=damp, damp_depth=damp_depth,
            damp_max=damp_max, damp_at_every_stage=damp_at_every_stage)
        self.diag = IsentropicDiagnostics(grid)
        self.spare = {n: storage.zeros(self.dyc.storage_shape, device=device) for n in self.out_names}
        self.nstep = 0

    def prepare_step(self):
        self.nstep += 1
        self.dyc.update_topography(self.nstep * self.dt)
        self.dyc._prognostic._diagnostics._set_topography()
        self.diag._set_topography()

    def compute_step(self):
        """dycore -> diagnostics refresh; ping-pong the output buffers."""
        out = self.dyc(self.state, {}, self.dt, out_state=self.spare)
        new = {n: out[n] for n in self.out_names}
        for n in (P, EXN, H, MTG):
            new[n] = self.state[n]
        self.spare = {n: self.state[n] for n in self.out_names}
        self.diag.get_diagnostic_variables(new[S], self.pt, new[P], new[EXN], new[MTG], new[H])
        self.state = new

    def finish_step(self):
        self.state["time"] = self.init_time + self.nstep * self.dt

    def buffer_dicts(self):
        return [self.state, self.spare]

    def set_buffer_dicts(self, dicts):
        self.state, self.spare = dicts

    def step(self):
        self.prepare_step()
        self.compute_step()
        self.finish_step()
        return self.state
