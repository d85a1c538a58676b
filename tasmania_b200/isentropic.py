# -*- coding: utf-8 -*-
"""The isentropic dynamical core on b200 storages -- host-side mirror of

  IsentropicDiagnostics    src/tasmania/isentropic/dynamics/diagnostics.py:L49-L597
  IsentropicPrognostic     src/tasmania/isentropic/dynamics/prognostic.py:L54-L174
    RK3WSSI                .../subclasses/prognostics/rk3ws_si.py:L37-L271
    ForwardEulerSI         .../subclasses/prognostics/forward_euler_si.py:L37-L233
  IsentropicDynamicalCore  src/tasmania/isentropic/dynamics/dycore.py:L55-L855

at the raw-array level (``stage_array_call`` and below; the DataArray/sympl layer above it is
orchestration that stays in tasmania).  Two execution paths produce the same numbers:

* the *stencil* path issues one launch per reference stencil, exactly as the reference does
  (works for every boundary type, moist or dry);
* the *fused* path (dry, relaxed boundary -- the benchmark configuration) runs a whole RK
  stage in two kernels through ``tb200_isentropic_stage_dry``.
"""
from __future__ import annotations

import ctypes as C
import os


import numpy as np

from tasmania_b200 import lib, storage
from tasmania_b200.dwarfs import HorizontalVelocity, VerticalDamping, WaterConstituent
from tasmania_b200.framework import BackendOptions, GridComponent, StencilFactory, StorageOptions
from tasmania_b200.grid import CONSTANTS
from tasmania_b200.stencils import FLUX

mfwv = "mass_fraction_of_water_vapor_in_air"
mfcw = "mass_fraction_of_cloud_liquid_water_in_air"
mfpw = "mass_fraction_of_precipitation_water_in_air"
S, SU, SV = "air_isentropic_density", "x_momentum_isentropic", "y_momentum_isentropic"
U, V = "x_velocity_at_u_locations", "y_velocity_at_v_locations"
MTG = "montgomery_potential"
SQV = "isentropic_density_of_water_vapor"
SQC = "isentropic_density_of_cloud_liquid_water"
SQR = "isentropic_density_of_precipitation_water"


class IsentropicDiagnostics(GridComponent, StencilFactory):
    """Pressure, Exner function, Montgomery potential and height of the interface levels by
    per-column vertical scans (K3)."""

    def __init__(self, grid, physical_constants=None, *, backend="b200", backend_options=None,
                 storage_shape=None, storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        self.grid = grid
        self.rpc = dict(CONSTANTS)
        # plain numbers under the externals' names (pref, rd, g, cp) or under the reference's long
        # names (diagnostics.py:L56-L63), in the reference's units
        for key, value in (physical_constants or {}).items():
            self.rpc[self.physical_constant_names.get(key, key)] = float(getattr(value, "values", value))
        nx, ny, nz = grid.nx, grid.ny, grid.nz
        shape = tuple(storage_shape or (nx + 1, ny + 1, nz + 1))
        self._shape = shape
        # theta is rank-1 in k: (1, 1, nk) storage seen through a zero-stride view
        th = np.zeros(shape[2])
        th[: nz + 1] = grid.z_on_interface_levels
        self._theta1d = storage.as_storage(th[None, None, :], device=self.storage_options.device)
        self._theta = storage.B200Array(self._theta1d.t.expand(shape[0], shape[1], -1))
        # topography lives at level nz of a 3-D storage in the reference
        # (diagnostics.py:L172-L174); a (ni, nj, 1) storage viewed with stride 0 along k here
        self._topo2d = self.zeros(shape=(shape[0], shape[1], 1))
        self._topo = storage.B200Array(self._topo2d.t.expand(-1, -1, shape[2]))
        self.backend_options.externals = dict(self.rpc)
        self._stencil_diagnostic_variables = self.compile_stencil("diagnostic_variables")
        self._stencil_density_and_temperature = self.compile_stencil("density_and_temperature")
        self._stencil_montgomery = self.compile_stencil("montgomery")
        self._stencil_height = self.compile_stencil("height")

    physical_constant_names = {
        "air_pressure_at_sea_level": "pref", "gas_constant_of_dry_air": "rd",
        "gravitational_acceleration": "g", "specific_heat_of_dry_air_at_constant_pressure": "cp"}
    default_physical_constants = {long: CONSTANTS[short] for long, short in physical_constant_names.items()}

    @property
    def raw_physical_constants(self):
        """framework/base_components.py:L46-L48, keyed by the reference's names."""
        return {long: self.rpc[short] for long, short in self.physical_constant_names.items()}

    def _set_topography(self):
        """Current terrain height -> device.  The reference re-uploads the host profile on
        every call (diagnostics.py:L172-L174); here the steady profile is uploaded once and
        ``profile = fact * steady`` (src/tasmania/domain/topography.py:L106-L116) is evaluated
        by the ``scale`` kernel whenever the growth factor changed -- the same single
        multiplication, hence the same bits, without the per-step H2D copy."""
        g, topo = self.grid, self.grid.topography
        fact = getattr(topo, "_fact", None)
        if fact is None:  # foreign topography object: plain upload
            self._topo2d[: g.nx, : g.ny, 0] = np.ascontiguousarray(topo.profile)
            return
        if getattr(self, "_steady_src", None) is not topo.steady_profile:
            self._steady_src = topo.steady_profile
            self._steady2d = self.zeros(shape=self._topo2d.shape)
            self._steady2d[: g.nx, : g.ny, 0] = np.ascontiguousarray(topo.steady_profile)
            self._fact_on_device = None
        if self._fact_on_device != fact:
            from tasmania_b200.stencils import _ew

            _ew("scale", self._topo2d, self._steady2d, f=fact, origin=(0, 0, 0),
                domain=(g.nx, g.ny, 1))
            self._fact_on_device = fact

    def get_diagnostic_variables(self, s, pt, p, exn, mtg, h):
        g = self.grid
        self._set_topography()
        self._stencil_diagnostic_variables(
            in_theta=self._theta, in_hs=self._topo, in_s=s, inout_p=p, out_exn=exn, inout_mtg=mtg,
            inout_h=h, dz=g.dz, pt=pt, origin=(0, 0, 0), domain=(g.nx, g.ny, g.nz + 1))

    def get_montgomery_potential(self, s, pt, mtg):
        g = self.grid
        self._set_topography()
        self._stencil_montgomery(
            in_hs=self._topo, in_s=s, inout_mtg=mtg, dz=g.dz, pt=pt,
            theta_s=float(g.z_on_interface_levels[-1]), origin=(0, 0, 0),
            domain=(g.nx, g.ny, g.nz + 1))

    def get_height(self, s, pt, h):
        g = self.grid
        self._set_topography()
        self._stencil_height(in_theta=self._theta, in_hs=self._topo, in_s=s, inout_h=h, dz=g.dz,
                             pt=pt, origin=(0, 0, 0), domain=(g.nx, g.ny, g.nz + 1))

    def get_density_and_temperature(self, s, exn, h, rho, t):
        g = self.grid
        self._stencil_density_and_temperature(
            in_theta=self._theta, in_s=s, in_exn=exn, in_h=h, out_rho=rho, out_t=t,
            origin=(0, 0, 0), domain=(g.nx, g.ny, g.nz))


class IsentropicPrognostic(GridComponent, StencilFactory):
    """Prognostic stage (K1 -> boundary(s) -> Montgomery -> K2); ``factory`` by scheme name."""

    name = None

    def __init__(self, horizontal_flux_scheme, grid, horizontal_boundary, moist, *, backend="b200",
                 backend_options=None, storage_shape=None, storage_options=None, pt=0.0, eps=0.5):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        self.grid, self.horizontal_boundary, self._moist = grid, horizontal_boundary, moist
        if horizontal_flux_scheme not in FLUX:
            raise ValueError(f"unknown horizontal flux scheme {horizontal_flux_scheme!r}")
        self._hflux = FLUX[horizontal_flux_scheme]
        g, hb = grid, horizontal_boundary
        self._storage_shape = tuple(storage_shape or (g.nx + 1, g.ny + 1, g.nz + 1))
        assert hb.nb >= self._hflux.extent, (
            f"The number of lateral boundary layers is {hb.nb}, but should be "
            f"greater or equal than {self._hflux.extent}.")
        assert g.nx >= 2 * hb.nb + 1 and g.ny >= 2 * hb.nb + 1
        self._pt = float(pt)
        self._eps = eps
        assert 0.0 <= eps <= 1.0, "The off-centering parameter should be between 0 and 1."
        self._diagnostics = IsentropicDiagnostics(
            g, backend=backend, backend_options=BackendOptions(), storage_shape=self._storage_shape,
            storage_options=self.storage_options)
        self._stencil = None
        self._stencil_momentum = None
        self._mtg_new = None

    @staticmethod
    def factory(scheme, *args, **kwargs):
        classes = {"rk3ws_si": RK3WSSI, "forward_euler_si": ForwardEulerSI}
        if scheme not in classes:
            raise ValueError(f"unknown time integration scheme {scheme!r}")
        return classes[scheme](*args, **kwargs)

    def _stencils_initialize(self, tendencies):
        externals = {
            "extent": self._hflux.extent,
            "flux_dry": self._hflux,
            "flux_moist": self._hflux,
            "moist": self._moist,
            "s_tnd_on": S in tendencies,
            "su_tnd_on": SU in tendencies,
            "sv_tnd_on": SV in tendencies,
            "qv_tnd_on": self._moist and mfwv in tendencies,
            "qc_tnd_on": self._moist and mfcw in tendencies,
            "qr_tnd_on": self._moist and mfpw in tendencies,
        }
        self.backend_options.externals = externals
        self._stencil = self.compile_stencil("step_forward_euler")
        self._stencil_momentum = self.compile_stencil("step_forward_euler_momentum")
        self._mtg_new = self.zeros(shape=self._storage_shape)

    def substep(self, stage, timestep):
        """(dtr, dt): the increment of the time label and the stage time step."""
        raise NotImplementedError

    def stage_call(self, stage, timestep, state, tendencies, out_state):
        g, nb = self.grid, self.horizontal_boundary.nb
        nx, ny, nz = g.nx, g.ny, g.nz
        tendencies = tendencies or {}
        if self._stencil is None:
            self._stencils_initialize(tendencies)
        dtr, dt = self.substep(stage, timestep)
        if stage == 0:
            self._now = {n: state[n] for n in (S, MTG, SU, SV)}
            if self._moist:
                self._now.update({n: state[n] for n in (SQV, SQC, SQR)})
        dt = dt.total_seconds()
        now = self._now
        args = dict(s_now=now[S], s_int=state[S], s_tnd=tendencies.get(S), s_new=out_state[S],
                    u_int=state[U], v_int=state[V], su_int=state[SU], sv_int=state[SV])
        if self._moist:
            args.update(
                sqv_now=now[SQV], sqv_int=state[SQV], qv_tnd=tendencies.get(mfwv), sqv_new=out_state[SQV],
                sqc_now=now[SQC], sqc_int=state[SQC], qc_tnd=tendencies.get(mfcw), sqc_new=out_state[SQC],
                sqr_now=now[SQR], sqr_int=state[SQR], qr_tnd=tendencies.get(mfpw), sqr_new=out_state[SQR])
        origin, domain = (nb, nb, 0), (nx - 2 * nb, ny - 2 * nb, nz)
        self._stencil(**args, dt=dt, dx=g.dx, dy=g.dy, origin=origin, domain=domain)
        self.horizontal_boundary.enforce_field(out_state[S], S, "kg m^-2 K^-1",
                                               time=state.get("time"))
        self._diagnostics.get_montgomery_potential(out_state[S], self._pt, self._mtg_new)
        self._stencil_momentum(
            s_now=now[S], s_int=state[S], s_new=out_state[S], u_int=state[U], v_int=state[V],
            mtg_now=now[MTG], mtg_new=self._mtg_new, su_now=now[SU], su_int=state[SU],
            su_tnd=tendencies.get(SU), su_new=out_state[SU], sv_now=now[SV], sv_int=state[SV],
            sv_tnd=tendencies.get(SV), sv_new=out_state[SV], dt=dt, dx=g.dx, dy=g.dy, eps=self._eps,
            origin=origin, domain=domain)
        if "time" in state:
            out_state["time"] = state["time"] + dtr


class ForwardEulerSI(IsentropicPrognostic):
    name = "forward_euler_si"
    stages = 1
    substep_fractions = 1.0

    def substep(self, stage, timestep):
        return timestep, timestep


class RK3WSSI(IsentropicPrognostic):
    name = "rk3ws_si"
    stages = 3
    substep_fractions = (1.0 / 3.0, 0.5, 1.0)

    def substep(self, stage, timestep):
        # timedelta arithmetic as in rk3ws_si.py:L115-L123 (microsecond rounding included)
        if stage == 0:
            return timestep / 3.0, timestep / 3.0
        if stage == 1:
            return timestep / 6.0, 0.5 * timestep
        return 0.5 * timestep, timestep


class IsentropicDynamicalCore(GridComponent, StencilFactory):
    """Raw-array dynamical core: ``stage_array_call`` (dry and moist) + the stage chaining of
    ``DynamicalCore.__call__`` (src/tasmania/framework/dycore.py:L383-L462)."""

    def __init__(self, grid, horizontal_boundary, *, moist=False,
                 time_integration_scheme="forward_euler_si", horizontal_flux_scheme="upwind",
                 time_integration_properties=None, damp=True, damp_at_every_stage=True,
                 damp_type="rayleigh", damp_depth=15, damp_max=0.0002, fused=None,
                 backend="b200", backend_options=None, storage_shape=None, storage_options=None,
                 smooth=True, smooth_at_every_stage=True, smooth_moist=False,
                 smooth_moist_at_every_stage=True, **smoothing_options):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        self.grid, self.horizontal_boundary = grid, horizontal_boundary
        self._moist, self._damp, self._damp_at_every_stage = moist, damp, damp_at_every_stage
        # the reference's constructor takes twelve smooth* arguments (dycore.py:L81-L92), keeps four
        # flags (L245-L248) and never applies any smoothing in its stages (L641-L843; the benchmark
        # drivers smooth through the IsentropicHorizontalSmoothing component): accepted and kept
        # likewise, so that a reference call site carries over; anything else is a mistake
        unknown = set(smoothing_options) - {
            "smooth_type", "smooth_coeff", "smooth_coeff_max", "smooth_damp_depth", "smooth_moist_type",
            "smooth_moist_coeff", "smooth_moist_coeff_max", "smooth_moist_damp_depth"}
        if unknown:
            raise TypeError(f"IsentropicDynamicalCore: unexpected keyword argument(s) {sorted(unknown)}")
        self._smooth, self._smooth_at_every_stage = smooth, smooth_at_every_stage
        self._smooth_moist, self._smooth_moist_at_every_stage = smooth_moist, smooth_moist_at_every_stage
        g = grid
        self.storage_shape = tuple(storage_shape or (g.nx + 1, g.ny + 1, g.nz + 1))
        kwargs = time_integration_properties or {}
        self._prognostic = IsentropicPrognostic.factory(
            time_integration_scheme, horizontal_flux_scheme, grid, horizontal_boundary, moist,
            backend=backend, backend_options=BackendOptions(), storage_shape=self.storage_shape,
            storage_options=self.storage_options, **kwargs)
        if damp:
            self._damper = VerticalDamping.factory(
                damp_type, grid, damp_depth, damp_max, backend=backend,
                backend_options=BackendOptions(), storage_shape=self.storage_shape,
                storage_options=self.storage_options)
        self._velocity_components = HorizontalVelocity(
            grid, staggering=True, backend=backend, backend_options=BackendOptions(),
            storage_options=self.storage_options)
        if moist:
            self._water_constituent = WaterConstituent(
                grid, clipping=True, backend=backend, backend_options=BackendOptions(),
                storage_options=self.storage_options)
            z = lambda: self.zeros(shape=self.storage_shape)  # noqa: E731
            self._sq = {(t, n): z() for t in ("now", "int", "new") for n in (SQV, SQC, SQR)}
        from tasmania_b200.boundary import Periodic, Relaxed

        # the 2-D class only: Relaxed1DX / 1DY also report type "relaxed", but their gamma is
        # non-zero on the middle line alone and they repeat that line across the degenerate axis,
        # which the fused kernels do not do
        fusable = type(horizontal_boundary) is Relaxed
        # ... and the 2-D Periodic class: same kernels with gamma = 0, the wrap of s
        # between the s-step and the scans inside the stage call, enforce_raw and damping after it
        self._periodic = (type(horizontal_boundary) is Periodic
                          and os.environ.get("TB200_FUSED_PERIODIC", "1") != "0"
                          and bool(lib.load().tb200_stage_lazy_velocities(grid.nz)))
        fusable = fusable or self._periodic
        if moist:  # the tracer kernel rides the default kernel path only (csrc/isentropic_fused.cu: kernel T)
            fusable = (fusable and os.environ.get("TB200_MOIST_FUSED", "1") != "0"
                       and bool(lib.load().tb200_stage_lazy_velocities(grid.nz)))
        if fused and not fusable:
            raise ValueError("the fused stage covers the core with (2-D) relaxed or periodic boundaries only")
        self._fused = fusable if fused is None else bool(fused)
        # fused path only: intermediate RK stages neither write nor read u, v (see _stage_fused);
        # set to False to get every stage's velocities like the reference's stage_array_call
        # (TB200_LAZY_UV=0 or an earlier kernel variant forced by the environment: the reference's flow)
        self.lazy_velocities = bool(self._fused and os.environ.get("TB200_LAZY_UV", "1") != "0"
                                    and lib.load().tb200_stage_lazy_velocities(grid.nz))
        # ... and stage 0 too re-diagnoses them instead of reading the state's u, v.  Only for
        # callers who know that those ARE the diagnosis of the state's s, su, sv (true for every
        # state this core produced, not for an initial state given in terms of u, v)
        self.derive_stage0_velocities = False
        self._raw_stage_states = None
        self._s_now = self._su_now = self._sv_now = None
        self._ref = None
        self._scratch = None
        if self._fused:
            self._allocate_scratch()  # now, not lazily: a raw device allocation must not fall into a graph capture
        # hook run after every stage on that stage's output fields (halo exchange of a
        # decomposed run); None on a single device
        self.after_stage = None
        # overlap controller of a decomposed run (tasmania_b200.distributed.Overlap) or None
        self.overlap = None

    @property
    def stages(self):
        return self._prognostic.stages

    @property
    def output_names(self):
        return (S, SU, U, SV, V) + ((mfwv, mfcw, mfpw) if self._moist else ())

    def allocate_stage_outputs(self):
        return {n: self.zeros(shape=self.storage_shape) for n in self.output_names}

    def update_topography(self, elapsed):
        self.grid.update_topography(elapsed)

    # ---- dycore.py:L641-L843
    def stage_array_call(self, stage, state, tendencies, timestep, out_state):
        hb = self.horizontal_boundary
        if stage == 0:
            self._s_now, self._su_now, self._sv_now = state[S], state[SU], state[SV]
            try:
                self._ref = {n: hb.reference_state[n] for n in (S, SU, SV)}
            except KeyError:
                if self._damp:
                    raise RuntimeError(
                        "Reference state not set in the object handling the horizontal boundary "
                        "conditions, but needed by the wave absorber.") from None
        slow = {n: v for n, v in (tendencies or {}).items() if n != "time"}
        if self._fused and slow and self._tendencies_fusable(slow):
            # slow tendencies of s, su, sv ride the fused stage (rk3ws_si.py:L105-L234 passes them to
            # K1 / K2); a missing one is a field of zeros: x - 0.0 == x, the reference's own form
            return self._stage_fused(stage, state, timestep, out_state, tendencies=slow)
        if self._fused and not slow:
            if self.overlap is None:
                return self._stage_fused(stage, state, timestep, out_state)
            # communication / computation overlap of a decomposed run: everything a neighbour
            # needs first, then the halo exchange (started by the hook on its own stream) runs
            # under the interior blocks of the momentum kernel
            self._stage_fused(stage, state, timestep, out_state, part=1, rim=self.overlap.rim)
            self.overlap.after_rim(stage, out_state)
            self._stage_fused(stage, state, timestep, out_state, part=2, rim=self.overlap.rim)
            self.overlap.after_interior(stage, out_state)
            return None
        if self._moist:
            wc = self._water_constituent
            tag = "now" if stage == 0 else "int"
            for qn, sqn in ((mfwv, SQV), (mfcw, SQC), (mfpw, SQR)):
                wc.get_density_of_water_constituent(state[S], state[qn], self._sq[(tag, sqn)])
                state[sqn] = self._sq[(tag, sqn)]
                out_state[sqn] = self._sq[("new", sqn)]
        self._prognostic.stage_call(stage, timestep, state, tendencies, out_state)
        if self._moist:
            for qn, sqn in ((mfwv, SQV), (mfcw, SQC), (mfpw, SQR)):
                wc.get_mass_fraction_of_water_constituent_in_air(
                    out_state[S], out_state.pop(sqn), out_state[qn])
        hb.enforce_raw(out_state, {n: {} for n in self.output_names})
        s_new, su_new, sv_new = out_state[S], out_state[SU], out_state[SV]
        if self._damp and (self._damp_at_every_stage or stage == self.stages - 1):
            self._damper(timestep, self._s_now, s_new, self._ref[S], s_new)
            self._damper(timestep, self._su_now, su_new, self._ref[SU], su_new)
            self._damper(timestep, self._sv_now, sv_new, self._ref[SV], sv_new)
        self._velocity_components.get_velocity_components(s_new, su_new, sv_new, out_state[U],
                                                          out_state[V])
        hb.set_outermost_layers_x(out_state[U], field_name=U, time=out_state.get("time"))
        hb.set_outermost_layers_y(out_state[V], field_name=V, time=out_state.get("time"))

    def _allocate_scratch(self):
        """The hand-off arrays of the fused stage (parked pressures, new Montgomery potential, s after
        its first relaxation) live in a library context held by this object (``tb200_ctx``, SURVEY.md
        section 8b): allocated once, freed with the dycore.  Host storages (the CPU test double of the
        library) keep plain storages."""
        ctx, self._scratch = storage.stage_scratch(self.storage_shape, 3, self.storage_options.device)
        if ctx is not None:
            self._ctx = ctx
        if self._periodic:
            self._periodic_gamma()  # allocated with the scratch, outside any graph capture

    def _periodic_gamma(self):
        if getattr(self, "_zero_gamma2d", None) is None:
            self._zero_gamma2d = self.zeros(shape=self.storage_shape[:2] + (1,))
        return self._zero_gamma2d

    def _tendencies_fusable(self, slow):
        return (not self._moist and self.overlap is None and set(slow) <= {S, SU, SV}
                and all(isinstance(v, storage.B200Array) and tuple(v.shape) == self.storage_shape
                        for v in slow.values())
                and bool(lib.load().tb200_stage_lazy_velocities(self.grid.nz))
                and os.environ.get("TB200_FUSED_TENDENCIES", "1") != "0")

    # ---- the fused stage: three kernels
    def _stage_fused(self, stage, state, timestep, out_state, part=0, rim=(0, 0, 0, 0), tendencies=None):
        g, hb, pr = self.grid, self.horizontal_boundary, self._prognostic
        qn = (mfwv, mfcw, mfpw) if self._moist else ()
        if stage == 0:
            pr._now = {n: state[n] for n in (S, MTG, SU, SV) + qn}
        if self._scratch is None:
            self._allocate_scratch()
        dtr, dt = pr.substep(stage, timestep)
        cfg = lib.StageCfg()
        cfg.nx, cfg.ny, cfg.nz, cfg.nb = g.nx, g.ny, g.nz, hb.nb
        cfg.flux_scheme = pr._hflux.code
        damp = self._damp and (self._damp_at_every_stage or stage == self.stages - 1)
        periodic = self._periodic
        # periodic: the wrap of su, sv (enforce_raw) comes between the momentum step and the damping
        # (dycore.py:L684-L700), so the damping cannot ride the momentum kernel
        cfg.damp = int(damp and not periodic)
        cfg.periodic = int(periodic)
        cfg.dt, cfg.dt_full = dt.total_seconds(), timestep.total_seconds()
        cfg.dx, cfg.dy, cfg.dz, cfg.eps = g.dx, g.dy, g.dz, pr._eps
        cfg.pt, cfg.theta_s = pr._pt, float(g.z_on_interface_levels[-1])
        rpc = pr._diagnostics.rpc
        cfg.constants[:] = [rpc["pref"], rpc["rd"], rpc["g"], rpc["cp"]]
        cfg.part = part
        cfg.rim[:] = list(rim)
        # velocities between the stages of one step: an intermediate stage does not write u, v,
        # its successor re-diagnoses them from s, su, sv (same formula, same bits: the reference
        # computes them exactly so at the end of every stage, dycore.py:L702-L721); the stage
        # outputs of intermediate stages therefore hold STALE u, v (``lazy_velocities``)
        # The last stage does not write them either: one pass of tb200_velocity_components over the
        # finished state (40 B/pt, ~0.5 ms at config 5) replaces the in-kernel diagnosis, which cost
        # the momentum kernel a recomputed halo lane, a warm-up row per strip and spilled registers
        # (2.23 ms against 1.36 ms without it, profiles/README.md round 2).
        lazy = self.lazy_velocities
        last = stage == self.stages - 1
        cfg.derive_uv_in = int(lazy and (stage > 0 or self.derive_stage0_velocities))
        # (a periodic stage never writes them: the wrap and the damping that follow the call come first)
        cfg.skip_uv_out = int(lazy or periodic)
        # ... and then nobody re-reads s before it is final: the stage updates it in place
        scratch_s = out_state[S] if (cfg.skip_uv_out and part == 0) else self._scratch[2]
        keep_tnd = None
        if tendencies:
            if getattr(self, "_zero_tnd", None) is None:
                self._zero_tnd = self.zeros(shape=self.storage_shape)
            keep_tnd = [lib.as_field(tendencies.get(n, self._zero_tnd)) for n in (S, SU, SV)]
            cfg.s_tnd, cfg.su_tnd, cfg.sv_tnd = (C.pointer(k) for k in keep_tnd)
        pr._diagnostics._set_topography()
        ref, now = hb.reference_state, pr._now
        gamma2d = hb._gamma2d if not periodic else self._periodic_gamma()
        if periodic:  # no relaxation, no in-kernel damping, no velocity output: never read
            ref = {n: ref.get(n, state[n]) for n in (S, SU, SV, U, V) + qn}
        f = lib.as_field
        rmat = self._damper._rmat if self._damp else None
        args = (cfg, f(now[S]), f(now[SU]), f(now[SV]), f(now[MTG]),
                f(state[S]), f(state[SU]), f(state[SV]), f(state[U]), f(state[V]),
                f(out_state[S]), f(out_state[SU]), f(out_state[SV]), f(out_state[U]), f(out_state[V]),
                f(ref[S]), f(ref[SU]), f(ref[SV]), f(ref[U]), f(ref[V]),
                f(gamma2d), f(rmat), f(pr._diagnostics._topo2d),
                f(self._scratch[0]), f(self._scratch[1]), f(scratch_s))
        if self._moist:
            # the three water constituents ride along (kernel T): mass fractions in, mass fractions out
            keep = [[lib.as_field(d[n]) for n in qn] for d in (now, state, out_state, ref)]
            arrays = [(lib.FieldP * 3)(*[C.pointer(k) for k in ks]) for ks in keep]
            rc = lib.load().tb200_isentropic_stage_moist(*args, *arrays, lib.current_stream())
            lib.check(rc, "tb200_isentropic_stage_moist")
            del keep
        else:
            rc = lib.load().tb200_isentropic_stage_dry(*args, lib.current_stream())
            lib.check(rc, "tb200_isentropic_stage_dry")
        if "time" in state and part != 1:
            out_state["time"] = state["time"] + dtr
        if periodic:  # dycore.py:L684-L700 in the reference's order
            hb.enforce_raw({n: out_state[n] for n in (S, SU, SV) + qn} | {"time": out_state.get("time")})
            if damp:
                self._damper(timestep, self._s_now, out_state[S], self._ref[S], out_state[S])
                self._damper(timestep, self._su_now, out_state[SU], self._ref[SU], out_state[SU])
                self._damper(timestep, self._sv_now, out_state[SV], self._ref[SV], out_state[SV])
        # a decomposed run diagnoses the velocities itself, after the halo exchange of s, su, sv
        if lazy and last and part != 1 and self.after_stage is None and self.overlap is None:
            self.diagnose_velocities(out_state)
        elif periodic and not lazy:  # every stage's velocities, like the reference's stage_array_call
            self.diagnose_velocities(out_state)

    def diagnose_velocities(self, out_state):
        """u, v of a finished state from its s, su, sv, outermost faces from the reference state
        (dwarfs/diagnostics.py:L219-L272 + relaxed.py:L161-L191) in one launch."""
        g, ref = self.grid, self.horizontal_boundary.reference_state
        if self._periodic:  # outermost faces by the wrap (periodic.py:L116-L122), not from a reference state
            hb = self.horizontal_boundary
            self._velocity_components.get_velocity_components(out_state[S], out_state[SU], out_state[SV],
                                                              out_state[U], out_state[V])
            hb.set_outermost_layers_x(out_state[U], field_name=U, time=out_state.get("time"))
            hb.set_outermost_layers_y(out_state[V], field_name=V, time=out_state.get("time"))
            return
        f = lib.as_field
        rc = lib.load().tb200_velocity_components(
            f(out_state[S]), f(out_state[SU]), f(out_state[SV]), f(out_state[U]), f(out_state[V]),
            f(ref[U]), f(ref[V]), g.nx, g.ny, g.nz, lib.current_stream())
        lib.check(rc, "tb200_velocity_components")

    # ---- framework/dycore.py:L383-L462
    def stages_iter(self, state, tendencies, timestep, out_state=None):
        """Generator over the stages of one time step: each ``next()`` runs one stage and
        yields ``(stage, stage_output_dict)``.  A domain-decomposed run advances all its
        sub-domains stage by stage and exchanges halos in between
        (tasmania_b200.distributed); ``__call__`` simply exhausts it."""
        if self._raw_stage_states is None:
            self._raw_stage_states = [self.allocate_stage_outputs() for _ in range(self.stages - 1)]
        out_state = out_state if out_state is not None else {}
        for n in self.output_names:
            if n not in out_state:
                out_state[n] = self.zeros(shape=self.storage_shape)
        outs = self._raw_stage_states + [out_state]
        cur = state
        for stage in range(self.stages):
            # each stage sees its own dict (the moist path adds sq* entries to it)
            self.stage_array_call(stage, dict(cur), tendencies or {}, timestep, outs[stage])
            if self.after_stage is not None:
                self.after_stage(stage, outs[stage])
            if stage == self.stages - 1 and "time" in state:
                out_state["time"] = state["time"] + timestep
            yield stage, outs[stage]
            cur = outs[stage]

    def __call__(self, state, tendencies, timestep, out_state=None):
        out = None
        for _, out in self.stages_iter(state, tendencies, timestep, out_state):
            pass
        return out
