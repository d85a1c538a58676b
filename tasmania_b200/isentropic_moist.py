# -*- coding: utf-8 -*-
"""The moist isentropic model of BASELINE.json configs[2] assembled on the b200 backend: dynamical
core + Kessler microphysics with sedimentation, coupled by sequential-update splitting, exactly
the component list, order and time integrators of the reference's benchmark driver
(drivers/benchmarking/isentropic_moist/driver_namelist_sus.py:L110-L512 with the parameters of
namelist_sus.py:L33-L141).  Everything between two calls of ``step`` stays on the device.
"""
from __future__ import annotations

from datetime import datetime

from tasmania_b200 import storage
from tasmania_b200.boundary import HorizontalBoundary
from tasmania_b200.coupling import (
    AirPotentialTemperatureToDiagnostic,
    AirPotentialTemperatureToTendency,
    ConcurrentCoupling,
    IsentropicDiagnostics,
    IsentropicHorizontalSmoothing,
    IsentropicVelocityComponents,
    SequentialUpdateSplitting,
    TimeIntegrationOptions,
    update_swap,
)
from tasmania_b200.isentropic import IsentropicDynamicalCore
from tasmania_b200.isentropic_physics import (
    IsentropicConservativeCoriolis,
    IsentropicImplicitVerticalAdvectionDiagnostic,
    IsentropicSmagorinsky,
    IsentropicVerticalAdvection,
)
from tasmania_b200.microphysics import (
    KesslerFallVelocity,
    KesslerMicrophysics,
    KesslerSaturationAdjustmentPrognostic,
    KesslerSedimentation,
    Precipitation,
)

W = "tendency_of_air_potential_temperature"

# namelist_sus.py:L33-L141 (mass fractions in g g^-1)
NAMELIST_SUS = dict(
    nb=3, nr=6, time_integration_scheme="rk3ws_si", eps=0.5, physics_time_integration_scheme="rk2",
    horizontal_flux_scheme="fifth_order_upwind", vertical_advection=True,
    implicit_vertical_advection=False, vertical_flux_scheme="third_order_upwind", damp=True,
    damp_depth=15, damp_max=0.0005, damp_at_every_stage=False, smooth_type="second_order",
    smooth_coeff=1.0, smooth_coeff_max=1.0, smooth_damp_depth=0, smooth_moist=True,
    smooth_moist_coeff=1.0, smooth_moist_coeff_max=1.0, smooth_moist_damp_depth=0,
    smagorinsky_constant=0.18, coriolis_parameter=None, sedimentation_flux_scheme="second_order_upwind",
    rain_evaporation=True, autoconversion_threshold=0.1e-3, autoconversion_rate=0.001,
    collection_rate=2.2, saturation_rate=0.025,
)


class IsentropicMoistSUS:
    """``state`` maps names to numpy arrays (tasmania_b200.grid.isentropic_state_from_brunt_vaisala
    with ``moist=True, precipitation=True``); it is uploaded once.  ``step()`` is one pass of
    the driver's time loop; ``self.state`` is the current device-resident state."""

    def __init__(self, grid, state, timestep, init_time=None, device=None, **overrides):
        nl = dict(NAMELIST_SUS)
        nl.update(overrides)
        self.nl, self.grid, self.dt = nl, grid, timestep
        nx, ny, nz, nb = grid.nx, grid.ny, grid.nz, nl["nb"]
        shape = (nx + 1, ny + 1, nz + 1)
        up = lambda a: storage.as_storage(a, device=device)  # noqa: E731
        self.state = {n: up(v) for n, v in state.items()}
        self.init_time = init_time or datetime(1992, 2, 20)
        self.state["time"] = self.init_time
        if W not in self.state:  # driver_namelist_sus.py:L125-L132
            self.state[W] = storage.zeros(shape, device=device)
        hb = HorizontalBoundary.factory("relaxed", nx, ny, nz, nb, nr=nl["nr"])
        hb.reference_state = {n: up(v) for n, v in state.items()}
        self.horizontal_boundary = hb
        pt = float(state["air_pressure_on_interface_levels"][0, 0, 0])
        self.dycore = IsentropicDynamicalCore(
            grid, hb, moist=True, time_integration_scheme=nl["time_integration_scheme"],
            horizontal_flux_scheme=nl["horizontal_flux_scheme"],
            time_integration_properties={"pt": pt, "eps": nl["eps"]}, damp=nl["damp"],
            damp_depth=nl["damp_depth"], damp_max=nl["damp_max"],
            damp_at_every_stage=nl["damp_at_every_stage"], storage_shape=shape)
        ptis = nl["physics_time_integration_scheme"]
        kw = dict(storage_shape=shape)
        tio = TimeIntegrationOptions
        args = [tio(IsentropicDiagnostics(grid, True, pt, **kw))]
        cf_kw = {} if nl["coriolis_parameter"] is None else {"coriolis_parameter": nl["coriolis_parameter"]}
        args.append(tio(IsentropicConservativeCoriolis(grid, nb, **cf_kw), scheme=ptis))
        args.append(tio(IsentropicHorizontalSmoothing(
            grid, nb, nl["smooth_type"], nl["smooth_coeff"], nl["smooth_coeff_max"],
            nl["smooth_damp_depth"], moist=nl["smooth_moist"],
            smooth_moist_coeff=nl["smooth_moist_coeff"],
            smooth_moist_coeff_max=nl["smooth_moist_coeff_max"],
            smooth_moist_damp_depth=nl["smooth_moist_damp_depth"], **kw)))
        args.append(tio(IsentropicSmagorinsky(grid, nb, nl["smagorinsky_constant"]), scheme=ptis))
        args.append(tio(IsentropicVelocityComponents(grid, hb, **kw)))
        t2d = AirPotentialTemperatureToDiagnostic(grid, **kw)
        d2t = AirPotentialTemperatureToTendency(grid, **kw)
        ke = KesslerMicrophysics(
            grid, air_pressure_on_interface_levels=True,
            tendency_of_air_potential_temperature_in_diagnostics=False,
            rain_evaporation=nl["rain_evaporation"],
            autoconversion_threshold=nl["autoconversion_threshold"],
            autoconversion_rate=nl["autoconversion_rate"], collection_rate=nl["collection_rate"], **kw)
        args.append(tio(ConcurrentCoupling(ke, t2d, execution_policy="serial"), scheme=ptis))
        sa = KesslerSaturationAdjustmentPrognostic(
            grid, air_pressure_on_interface_levels=True, saturation_rate=nl["saturation_rate"], **kw)
        args.append(tio(ConcurrentCoupling(d2t, sa, t2d, execution_policy="serial"), scheme=ptis))
        if nl["vertical_advection"]:
            if nl["implicit_vertical_advection"]:
                args.append(tio(IsentropicImplicitVerticalAdvectionDiagnostic(grid, moist=True, **kw)))
            else:
                vf = IsentropicVerticalAdvection(grid, flux_scheme=nl["vertical_flux_scheme"],
                                                 moist=True, **kw)
                args.append(tio(vf, scheme="rk3ws"))
        rfv = KesslerFallVelocity(grid, **kw)
        sd = KesslerSedimentation(grid, sedimentation_flux_scheme=nl["sedimentation_flux_scheme"], **kw)
        args.append(tio(ConcurrentCoupling(rfv, sd, execution_policy="serial"), scheme="rk3ws"))
        ap = Precipitation(grid, **kw)
        args.append(tio(ConcurrentCoupling(rfv, ap, execution_policy="serial")))
        self.physics = SequentialUpdateSplitting(*args)
        self._state_new = None
        self.nstep = 0

    # ---- one step = prepare_step (host-dependent, tiny) + compute_step (a fixed kernel sequence
    # on fixed buffers: what tasmania_b200.graphs.GraphedLoop captures) + finish_step (host)
    def topography_consumers(self):
        """Every IsentropicDiagnostics core holding a device copy of the terrain height."""
        return [self.dycore._prognostic._diagnostics, self.physics._component_list[0]._core]

    def prepare_step(self):
        self.nstep += 1
        self.dycore.update_topography(self.nstep * self.dt)
        for d in self.topography_consumers():
            d._set_topography()

    def compute_step(self):
        """driver_namelist_sus.py:L490-L512."""
        state, dt = self.state, self.dt
        if self._state_new is None:
            self._state_new = {}
        out = self.dycore(state, {}, dt, out_state=self._state_new)
        # the variables the dycore does not return are carried over (update_swap, L503)
        update_swap(out, {n: state[n] for n in list(state) if n not in out and n != "time"})
        out["time"] = state["time"]
        self.physics(out, dt)
        # the arrays left in the old dict are the output buffers of the next step (L494)
        self.state, self._state_new = out, {n: v for n, v in state.items()
                                            if n in self.dycore.output_names}

    def finish_step(self):
        self.state["time"] = self.init_time + self.nstep * self.dt

    def buffer_dicts(self):
        """The dicts whose arrays are permuted by a step (update_swap / ping-pong)."""
        if self._state_new is None:
            self._state_new = {}
        return ([self.state, self._state_new] + self.physics._out_diagnostics
                + self.physics._out_state)

    def set_buffer_dicts(self, dicts):
        self.state, self._state_new = dicts[0], dicts[1]
        n = len(self.physics._out_diagnostics)
        self.physics._out_diagnostics[:] = dicts[2:2 + n]
        self.physics._out_state[:] = dicts[2 + n:]

    def step(self):
        self.prepare_step()
        self.compute_step()
        self.finish_step()
        return self.state

    def run(self, nsteps):
        for _ in range(nsteps):
            self.step()
        return self.state


def moist_mountain_case(nx, ny, nz, *, topo_seconds=60.0, max_height=1000.0, relative_humidity=0.98,
                        seed=True, half_width_km=(176.0, 176.0)):
    """BASELINE config 3, scalable: the moist mountain-flow case of namelist_sus.py
    (drivers/benchmarking/isentropic_moist/namelist_sus.py:L33-L141) on a ``2 half_width_km`` wide
    domain (a 161-point axis over 352 km keeps the benchmark's 2.2 km spacing), by default with a
    faster-growing, taller mountain and -- when ``seed`` -- blobs of cloud water and rain in the
    initial state so that autoconversion, accretion, evaporation, sedimentation and precipitation
    are all active within a few steps.  Returns (Grid, numpy state)."""
    from datetime import timedelta

    import numpy as np

    from tasmania_b200.grid import Grid, Topography, gaussian_profile, isentropic_state_from_brunt_vaisala
    from tasmania_b200.isentropic import mfcw, mfpw

    hx, hy = half_width_km
    x = np.linspace(-hx, hx, nx)
    y = np.linspace(-hy, hy, ny)
    topo = Topography(gaussian_profile(x, y, max_height, 50.0, 50.0), timedelta(seconds=topo_seconds))
    grid = Grid((-hx, hx), nx, (-hy, hy), ny, (400.0, 280.0), nz, units_to_m=1e3, topography=topo)
    state = isentropic_state_from_brunt_vaisala(grid, 22.5, 0.0, 0.015, moist=True, precipitation=True,
                                                relative_humidity=relative_humidity)
    if seed:
        i, j, k = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
        blob = np.exp(-(((i - nx // 2) / (0.2 * nx)) ** 2 + ((j - ny // 2) / (0.2 * ny)) ** 2
                        + ((k - 0.7 * nz) / (0.2 * nz)) ** 2))
        blob[nx:, :, :] = blob[:, ny:, :] = blob[:, :, nz:] = 0.0
        state[mfcw] = 8e-4 * blob
        state[mfpw] = 3e-4 * np.roll(blob, 2, axis=0) * (blob > 0)
    return grid, state
