# -*- coding: utf-8 -*-
"""Host-side mirrors of the isentropic physics components of SURVEY.md section 8f at the
raw-array level -- ``tasmania.IsentropicVerticalAdvection``
(src/tasmania/isentropic/physics/vertical_advection.py:L71-L269, row 1) and
``tasmania.IsentropicConservativeCoriolis`` (src/tasmania/isentropic/physics/coriolis.py:L44-L186,
row 3) and the Smagorinsky turbulence components ``tasmania.Smagorinsky2d`` /
``tasmania.IsentropicSmagorinsky`` (src/tasmania/physics/turbulence.py:L42-L229,
src/tasmania/isentropic/physics/turbulence.py:L38-L125, row 3): same constructor arguments, same externals, same ``array_call`` keyword wiring; the
arithmetic runs in ``tb200_vertical_advection`` (csrc/vertical.cu) and ``tb200_coriolis``
(csrc/elementwise.cu), ``tb200_smagorinsky`` (csrc/turbulence.cu) and, for
``tasmania.IsentropicImplicitVerticalAdvectionDiagnostic``
(src/tasmania/isentropic/physics/implicit_vertical_advection.py:L44-L336, row 4),
``tb200_implicit_vertical_advection`` (csrc/vertical.cu)."""
from __future__ import annotations

from tasmania_b200.framework import BackendOptions, GridComponent, StencilFactory, StorageOptions
from tasmania_b200.stencils import FLUX

S, SU, SV = "air_isentropic_density", "x_momentum_isentropic", "y_momentum_isentropic"
MFWV = "mass_fraction_of_water_vapor_in_air"
MFCW = "mass_fraction_of_cloud_liquid_water_in_air"
MFPW = "mass_fraction_of_precipitation_water_in_air"
W_ML = "tendency_of_air_potential_temperature"
W_HL = "tendency_of_air_potential_temperature_on_interface_levels"


class IsentropicVerticalAdvection(GridComponent, StencilFactory):
    """Vertical derivative of the conservative vertical advection flux of s, su, sv (and of the
    water species when ``moist``), by one of the minimal vertical flux schemes
    (src/tasmania/isentropic/dynamics/subclasses/minimal_vertical_fluxes/*.py)."""

    class_stencils = {"stencil": "vertical_advection"}

    def __init__(self, grid, flux_scheme="upwind", moist=False,
                 tendency_of_air_potential_temperature_on_interface_levels=False, *, backend="b200",
                 backend_options=None, storage_shape=None, storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        if flux_scheme not in FLUX:
            raise ValueError(f"unknown vertical flux scheme {flux_scheme!r}")
        self.grid = grid
        self._moist = moist
        self._stgz = tendency_of_air_potential_temperature_on_interface_levels
        self._vflux = FLUX[flux_scheme]
        # vertical_advection.py:L148-L160
        assert grid.nz >= 2 * self._vflux.extent + 1, "too few vertical levels for the flux scheme"
        self.storage_shape = tuple(storage_shape or (grid.nx + 1, grid.ny + 1, grid.nz + 1))
        self.backend_options.externals = {
            "flux_end": -self._vflux.extent + 1 if self._vflux.extent > 1 else None,
            "flux_extent": self._vflux.extent,
            "get_flux_dry": self._vflux,
            "get_flux_moist": self._vflux,
            "moist": moist,
            "set_output": self.get_subroutine_definition("set_output"),
            "staggering": self._stgz,
        }
        self._stencil = self.compile_stencil("stencil")
        self._stencil_step = self.compile_stencil("vertical_advection_step")
        self.tendency_names = (S, SU, SV) + ((MFWV, MFCW, MFPW) if moist else ())

    kind, diagnostic_names = "tendency", ()

    def array_call(self, state, out_tendencies, out_diagnostics=None, overwrite_tendencies=None):
        """vertical_advection.py:L216-L269"""
        g = self.grid
        ow = overwrite_tendencies or {}
        args = {
            "dz": g.dz,
            "in_w": state[W_HL] if self._stgz else state[W_ML],
            "in_s": state[S], "out_s": out_tendencies[S], "ow_out_s": ow.get(S, True),
            "in_su": state[SU], "out_su": out_tendencies[SU], "ow_out_su": ow.get(SU, True),
            "in_sv": state[SV], "out_sv": out_tendencies[SV], "ow_out_sv": ow.get(SV, True),
        }
        if self._moist:
            for key, name in (("qv", MFWV), ("qc", MFCW), ("qr", MFPW)):
                args["in_" + key] = state[name]
                args["out_" + key] = out_tendencies[name]
                args["ow_out_" + key] = ow.get(name, True)
        self._stencil(**args, origin=(0, 0, 0), domain=(g.nx, g.ny, g.nz))


    def array_call_stepped(self, state, base, factor, out_state):
        """b200 only -- one stage of a tendency stepper in one kernel: out_state[n] = base[n] +
        factor * tendency[n](state) for the advected fields (tasmania_b200.coupling.TendencyStepper
        uses it instead of ``array_call`` + ``fma``)."""
        g, names = self.grid, self.tendency_names
        self._stencil_step(in_w=state[W_HL] if self._stgz else state[W_ML],
                           ins=[state[n] for n in names], bases=[base[n] for n in names],
                           outs=[out_state[n] for n in names], dz=g.dz, factor=factor,
                           origin=(0, 0, 0), domain=(g.nx, g.ny, g.nz))


class IsentropicConservativeCoriolis(GridComponent, StencilFactory):
    """Mirror of ``tasmania.IsentropicConservativeCoriolis``
    (src/tasmania/isentropic/physics/coriolis.py:L44-L164): Coriolis forcing of the momenta on
    the interior of the numerical grid (``nb`` boundary layers excluded)."""

    def __init__(self, grid, nb, coriolis_parameter=1e-4, *, backend="b200", backend_options=None,
                 storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        self.grid, self._nb, self._f = grid, int(nb), float(coriolis_parameter)
        self.backend_options.externals = {"set_output": self.get_subroutine_definition("set_output")}
        self._stencil = self.compile_stencil("coriolis")

    kind, tendency_names, diagnostic_names = "tendency", (SU, SV), ()

    def array_call(self, state, out_tendencies, out_diagnostics=None, overwrite_tendencies=None):
        g, nb = self.grid, self._nb
        ow = overwrite_tendencies or {}
        self._stencil(in_su=state[SU], in_sv=state[SV], tnd_su=out_tendencies[SU],
                      tnd_sv=out_tendencies[SV], f=self._f, ow_tnd_su=ow.get(SU, True),
                      ow_tnd_sv=ow.get(SV, True), origin=(nb, nb, 0),
                      domain=(g.nx - 2 * nb, g.ny - 2 * nb, g.nz))

    def array_call_stepped(self, state, base, factor, out_state):
        """b200 only -- one stage of a tendency stepper in one kernel: out_state = base + factor *
        tendency(state) over the whole storages (tasmania_b200.coupling.TendencyStepper uses it
        instead of ``array_call`` + ``fma``)."""
        g, nb = self.grid, self._nb
        if not hasattr(self, "_stencil_step"):
            self._stencil_step = self.compile_stencil("coriolis_step")
        self._stencil_step(in_su=state[SU], in_sv=state[SV], base_su=base[SU], base_sv=base[SV],
                           out_su=out_state[SU], out_sv=out_state[SV], f=self._f, factor=factor,
                           origin=(nb, nb, 0), domain=(g.nx - 2 * nb, g.ny - 2 * nb, g.nz))


class Smagorinsky2d(GridComponent, StencilFactory):
    """Mirror of ``tasmania.Smagorinsky2d`` (src/tasmania/physics/turbulence.py:L42-L163):
    tendencies of x_velocity / y_velocity on the interior of the numerical grid."""

    names = ("x_velocity", "y_velocity")
    kind, diagnostic_names = "tendency", ()

    @property
    def tendency_names(self):
        return self.names

    def __init__(self, grid, nb, smagorinsky_constant=0.18, *, backend="b200", backend_options=None,
                 storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        assert nb >= 2, "The number of boundary layers must be greater or equal than two."
        self.grid, self._nb, self._cs = grid, max(2, int(nb)), float(smagorinsky_constant)
        self.backend_options.externals = {
            "core": self.get_subroutine_definition("smagorinsky_core"),
            "set_output": self.get_subroutine_definition("set_output"),
        }
        self._stencil = self.compile_stencil("smagorinsky")

    def _box(self):
        g, nb = self.grid, self._nb
        return dict(origin=(nb, nb, 0), domain=(g.nx - 2 * nb, g.ny - 2 * nb, g.nz))

    def array_call(self, state, out_tendencies, out_diagnostics=None, overwrite_tendencies=None):
        g, ow = self.grid, overwrite_tendencies or {}
        nu, nv = self.names
        self._stencil(in_u=state[nu], in_v=state[nv], out_u_tnd=out_tendencies[nu],
                      out_v_tnd=out_tendencies[nv], dx=g.dx, dy=g.dy, cs=self._cs,
                      ow_out_u_tnd=ow.get(nu, True), ow_out_v_tnd=ow.get(nv, True), **self._box())


class IsentropicSmagorinsky(Smagorinsky2d):
    """Mirror of ``tasmania.IsentropicSmagorinsky``
    (src/tasmania/isentropic/physics/turbulence.py:L38-L125): conservative form, tendencies of
    the momenta."""

    class_stencils = {"smagorinsky": "smagorinsky_isentropic"}
    names = (SU, SV)

    def array_call(self, state, out_tendencies, out_diagnostics=None, overwrite_tendencies=None):
        g, ow = self.grid, overwrite_tendencies or {}
        self._stencil(in_s=state[S], in_su=state[SU], in_sv=state[SV], out_su_tnd=out_tendencies[SU],
                      out_sv_tnd=out_tendencies[SV], dx=g.dx, dy=g.dy, cs=self._cs,
                      ow_out_su_tnd=ow.get(SU, True), ow_out_sv_tnd=ow.get(SV, True), **self._box())

    def array_call_stepped(self, state, base, factor, out_state):
        """b200 only -- one stage of a tendency stepper without the round trip of the tendencies
        through memory: out_state = base + factor * tendency(state) over the whole storages."""
        g = self.grid
        if not hasattr(self, "_stencil_step"):
            self._stencil_step = self.compile_stencil("smagorinsky_isentropic_step")
        self._stencil_step(in_s=state[S], in_su=state[SU], in_sv=state[SV], base_su=base[SU],
                           base_sv=base[SV], out_su=out_state[SU], out_sv=out_state[SV], dx=g.dx, dy=g.dy,
                           cs=self._cs, factor=factor, **self._box())


class IsentropicImplicitVerticalAdvectionDiagnostic(GridComponent, StencilFactory):
    """Mirror of ``tasmania.IsentropicImplicitVerticalAdvectionDiagnostic``
    (src/tasmania/isentropic/physics/implicit_vertical_advection.py:L44-L219): Crank-Nicolson
    vertical advection, one tridiagonal solve per column and field."""

    kind, tendency_names = "implicit", ()

    def diagnostic_shape(self, name):
        return self.storage_shape

    def __init__(self, grid, moist=False,
                 tendency_of_air_potential_temperature_on_interface_levels=False, *, backend="b200",
                 backend_options=None, storage_shape=None, storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        self.grid, self._moist = grid, moist
        self.storage_shape = tuple(storage_shape or (grid.nx + 1, grid.ny + 1, grid.nz + 1))
        self.diagnostic_names = (S, SU, SV) + ((MFWV, MFCW, MFPW) if moist else ())
        self._stgz = tendency_of_air_potential_temperature_on_interface_levels
        self.backend_options.externals = {  # L112-L117
            "moist": moist,
            "staggering": self._stgz,
            "setup": self.get_subroutine_definition("setup_thomas"),
            "setup_bc": self.get_subroutine_definition("setup_thomas_bc"),
        }
        self._stencil = self.compile_stencil("implicit_vertical_advection")

    def array_call(self, state, timestep, out_tendencies, out_diagnostics, overwrite_tendencies=None):
        """implicit_vertical_advection.py:L172-L219"""
        g = self.grid
        args = {
            "gamma": timestep.total_seconds() / (4.0 * g.dz),
            "in_w": state[W_HL] if self._stgz else state[W_ML],
            "in_s": state[S], "out_s": out_diagnostics[S],
            "in_su": state[SU], "out_su": out_diagnostics[SU],
            "in_sv": state[SV], "out_sv": out_diagnostics[SV],
        }
        if self._moist:
            for key, name in (("qv", MFWV), ("qc", MFCW), ("qr", MFPW)):
                args["in_" + key] = state[name]
                args["out_" + key] = out_diagnostics[name]
        self._stencil(**args, origin=(0, 0, 0), domain=(g.nx, g.ny, g.nz))


class IsentropicImplicitVerticalAdvectionPrognostic(GridComponent, StencilFactory):
    """Mirror of ``tasmania.IsentropicImplicitVerticalAdvectionPrognostic``
    (src/tasmania/isentropic/physics/implicit_vertical_advection.py:L593-L919): the same solves,
    returned as tendencies (x_new - x) / dt in storages owned by the component."""

    class_stencils = {"stencil": "implicit_vertical_advection_tendency"}

    def __init__(self, grid, moist=False,
                 tendency_of_air_potential_temperature_on_interface_levels=False, *, backend="b200",
                 backend_options=None, storage_shape=None, storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        self.grid, self._moist = grid, moist
        self._stgz = tendency_of_air_potential_temperature_on_interface_levels
        shape = tuple(storage_shape or (grid.nx + 1, grid.ny + 1, grid.nz + 1))
        names = (S, SU, SV) + ((MFWV, MFCW, MFPW) if moist else ())
        self._tnd = {n: self.zeros(shape=shape) for n in names}
        self.backend_options.externals = {  # L665-L670
            "moist": moist,
            "vstaggering": self._stgz,
            "setup": self.get_subroutine_definition("setup_thomas"),
            "setup_bc": self.get_subroutine_definition("setup_thomas_bc"),
        }
        self._stencil = self.compile_stencil("stencil")

    def array_call(self, state, timestep):
        """L722-L792: returns (tendencies, diagnostics)"""
        g = self.grid
        dt = timestep.total_seconds()
        args = {"dt": dt, "gamma": dt / (4.0 * g.dz), "in_w": state[W_HL] if self._stgz else state[W_ML]}
        for key, name in (("s", S), ("su", SU), ("sv", SV)) + (
                (("qv", MFWV), ("qc", MFCW), ("qr", MFPW)) if self._moist else ()):
            args["in_" + key] = state[name]
            args["tnd_" + key] = self._tnd[name]
        self._stencil(**args, origin=(0, 0, 0), domain=(g.nx, g.ny, g.nz))
        return dict(self._tnd), {}
