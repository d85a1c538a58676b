# -*- coding: utf-8 -*-
"""The 2-D Burgers dwarf on b200 storages (row K10 of SURVEY.md section 8a) -- host-side mirror,
at the raw-array level, of

  BurgersStepper (forward_euler, rk2, rk3ws)  src/tasmania/burgers/dynamics/stepper.py:L45-L335,
                                              .../subclasses/stepper/{forward_euler,rk2,rk3ws}.py
  BurgersDynamicalCore                        src/tasmania/burgers/dynamics/dycore.py:L38-L185
  ZhaoSolutionFactory                         src/tasmania/burgers/state.py:L42-L152
  BurgersHorizontalDiffusion                  src/tasmania/burgers/physics/diffusion.py:L40-L150

One ``forward_euler`` launch per stage steps u and v together (advection order 1..6 is a
template parameter of the kernel).  With Dirichlet boundaries the reference evaluates the analytic
rim values on the host on every stage and uploads them (dirichlet.py:L98-L150); the mirror keeps
that behaviour (``boundary.Dirichlet``), so configuration 1 is launch- and PCIe-latency bound.
"""
from __future__ import annotations

import numpy as np

from tasmania_b200.dwarfs import HorizontalDiffusion
from tasmania_b200.framework import BackendOptions, GridComponent, StencilFactory, StorageOptions
from tasmania_b200.stencils import ADVECTION


class ZhaoSolutionFactory:
    """Analytic solution of the viscous Burgers equations (Zhao et al.); its call signature is
    the Dirichlet ``core`` protocol (state.py:L59-L67, dirichlet.py:L57-L68)."""

    def __init__(self, initial_time, eps):
        self._itime, self._eps = initial_time, float(eps)

    def __call__(self, time, grid, slice_x=None, slice_y=None, field_name="x_velocity",
                 field_units=None):
        eps = self._eps
        sx = slice_x if slice_x is not None else slice(0, grid.nx)
        sy = slice_y if slice_y is not None else slice(0, grid.ny)
        x, y = grid.x[sx], grid.y[sy]
        mi, mj = len(x), len(y)
        x = np.tile(x[:, np.newaxis, np.newaxis], (1, mj, grid.nz))
        y = np.tile(y[np.newaxis, :, np.newaxis], (mi, 1, grid.nz))
        t = (time - self._itime).total_seconds()
        if field_name == "x_velocity":
            return (
                -2.0 * eps * 2.0 * np.pi * np.exp(-5.0 * np.pi**2 * eps * t)
                * np.cos(2.0 * np.pi * x) * np.sin(np.pi * y)
                / (2.0 + np.exp(-5.0 * np.pi**2 * eps * t) * np.sin(2.0 * np.pi * x) * np.sin(np.pi * y))
            )
        if field_name == "y_velocity":
            return (
                -2.0 * eps * np.pi * np.exp(-5.0 * np.pi**2 * eps * t)
                * np.sin(2.0 * np.pi * x) * np.cos(np.pi * y)
                / (2.0 + np.exp(-5.0 * np.pi**2 * eps * t) * np.sin(2.0 * np.pi * x) * np.sin(np.pi * y))
            )
        raise ValueError(f"unknown field {field_name!r}")


class BurgersStepper(StencilFactory):
    """Forward-Euler stages of the advection step; ``factory`` by scheme name."""

    SUBSTEPS = {
        # scheme: per stage (fraction of the time-label increment, fraction of dt)
        "forward_euler": ((1.0, 1.0),),
        "rk2": ((0.5, 0.5), (0.5, 1.0)),
        "rk3ws": ((1.0 / 3.0, 1.0 / 3.0), (1.0 / 6.0, 0.5), (0.5, 1.0)),
    }

    def __init__(self, scheme, grid, nb, flux_scheme, *, backend="b200", backend_options=None,
                 storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        if scheme not in self.SUBSTEPS:
            raise ValueError(f"unknown time integration scheme {scheme!r}")
        if flux_scheme not in ADVECTION:
            raise ValueError(f"unknown advection scheme {flux_scheme!r}")
        self.scheme, self._grid, self._nb = scheme, grid, nb
        self._advection = ADVECTION[flux_scheme]
        assert nb >= self._advection.extent
        self._forward_euler = None
        self._stencil_args = {}

    @staticmethod
    def factory(scheme, *args, **kwargs):
        return BurgersStepper(scheme, *args, **kwargs)

    @property
    def stages(self):
        return len(self.SUBSTEPS[self.scheme])

    def _stencil_initialize(self, tendencies):
        self.backend_options.externals = {
            "advection": self._advection,
            "extent": self._advection.extent,
            "tnd_u": "x_velocity" in tendencies,
            "tnd_v": "y_velocity" in tendencies,
        }
        self._forward_euler = self.compile_stencil("forward_euler")

    def __call__(self, stage, state, tendencies, timestep, out_state):
        g, nb = self._grid, self._nb
        if self._forward_euler is None:
            self._stencil_initialize(tendencies)
        fr, fdt = self.SUBSTEPS[self.scheme][stage]
        # timedelta / float arithmetic as in rk3ws.py:L48-L58
        dtr, dt = fr * timestep, fdt * timestep.total_seconds()
        if stage == 0:
            self._stencil_args["in_u"] = state["x_velocity"]
            self._stencil_args["in_v"] = state["y_velocity"]
        args = dict(self._stencil_args, in_u_tmp=state["x_velocity"], in_v_tmp=state["y_velocity"],
                    out_u=out_state["x_velocity"], out_v=out_state["y_velocity"])
        if "x_velocity" in tendencies:
            args["in_u_tnd"] = tendencies["x_velocity"]
        if "y_velocity" in tendencies:
            args["in_v_tnd"] = tendencies["y_velocity"]
        self._forward_euler(**args, dt=dt, dx=g.dx, dy=g.dy, origin=(nb, nb, 0),
                            domain=(g.nx - 2 * nb, g.ny - 2 * nb, 1))
        out_state["time"] = state["time"] + dtr


class BurgersDynamicalCore(GridComponent, StencilFactory):
    """Stage = stepper + lateral boundary (dycore.py:L158-L173), chained as in
    ``DynamicalCore.__call__`` (src/tasmania/framework/dycore.py:L383-L462)."""

    def __init__(self, grid, horizontal_boundary, time_integration_scheme="forward_euler",
                 flux_scheme="upwind", *, backend="b200", backend_options=None, storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        assert grid.nz == 1, "The number grid points along the vertical dimension must be 1."
        self.grid, self.horizontal_boundary = grid, horizontal_boundary
        self._stepper = BurgersStepper.factory(
            time_integration_scheme, grid, horizontal_boundary.nb, flux_scheme, backend=backend,
            backend_options=BackendOptions(), storage_options=self.storage_options)
        self._stage_states = None

    @property
    def stages(self):
        return self._stepper.stages

    def stage_array_call(self, stage, state, tendencies, timestep, out_state):
        self._stepper(stage, state, tendencies, timestep, out_state)
        self.horizontal_boundary.enforce_raw(
            out_state, {"x_velocity": {"units": "m s^-1"}, "y_velocity": {"units": "m s^-1"}})

    def __call__(self, state, tendencies, timestep, out_state=None):
        shape = state["x_velocity"].shape
        if self._stage_states is None:
            self._stage_states = [{n: self.zeros(shape=shape) for n in ("x_velocity", "y_velocity")}
                                  for _ in range(self.stages - 1)]
        out_state = out_state if out_state is not None else {}
        for n in ("x_velocity", "y_velocity"):
            if n not in out_state:
                out_state[n] = self.zeros(shape=shape)
        outs = self._stage_states + [out_state]
        cur = state
        for stage in range(self.stages):
            self.stage_array_call(stage, cur, tendencies or {}, timestep, outs[stage])
            cur = outs[stage]
        out_state["time"] = state["time"] + timestep
        return out_state


class BurgersHorizontalDiffusion(StencilFactory):
    """Diffusive tendencies of u and v (burgers/physics/diffusion.py:L134-L150)."""

    def __init__(self, grid, diffusion_type, diffusion_coeff, nb=None, *, backend="b200",
                 backend_options=None, storage_shape=None, storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        shape = tuple(storage_shape or (grid.nx, grid.ny, grid.nz))
        self._diffuser = HorizontalDiffusion.factory(
            diffusion_type, shape, grid.dx, grid.dy, diffusion_coeff, diffusion_coeff, 0, nb,
            backend=backend, backend_options=BackendOptions(), storage_options=self.storage_options)

    def array_call(self, state, out_tendencies, out_diagnostics, overwrite_tendencies):
        self._diffuser(state["x_velocity"], out_tendencies["x_velocity"],
                       overwrite_output=overwrite_tendencies["x_velocity"])
        self._diffuser(state["y_velocity"], out_tendencies["y_velocity"],
                       overwrite_output=overwrite_tendencies["y_velocity"])
