# -*- coding: utf-8 -*-
"""Kessler microphysics on b200 storages (row K11 of SURVEY.md section 8a) -- host-side mirror,
at the raw-array (``array_call``) level, of

  KesslerMicrophysics                    src/tasmania/physics/microphysics/kessler.py:L54-L452
  KesslerSaturationAdjustmentDiagnostic  kessler.py:L454-L772
  KesslerSaturationAdjustmentPrognostic  kessler.py:L774-L1088
  KesslerFallVelocity                    kessler.py:L1090-L1219
  KesslerSedimentation                   kessler.py:L1221-L1392
  Precipitation                          src/tasmania/physics/microphysics/utils.py:L144-L323

with the reference's constructor arguments (minus the sympl ``Domain``: a grid is enough at
this level), externals and stencil keyword names.  The DataArray / units / sympl component
layer above ``array_call`` is orchestration that stays in tasmania.
"""
from __future__ import annotations

import numpy as np

from tasmania_b200.framework import BackendOptions, GridComponent, StencilFactory, StorageOptions
from tasmania_b200.stencils import SEDIMENTATION_FLUX

mfwv = "mass_fraction_of_water_vapor_in_air"
mfcw = "mass_fraction_of_cloud_liquid_water_in_air"
mfpw = "mass_fraction_of_precipitation_water_in_air"

# kessler.py:L78-L82, L480-L486; utils.py:L147-L149
DEFAULT_CONSTANTS = {
    "gas_constant_of_dry_air": 287.05,
    "gas_constant_of_water_vapor": 461.52,
    "latent_heat_of_vaporization_of_water": 2.5e6,
    "specific_heat_of_dry_air_at_constant_pressure": 1004.0,
    "density_of_liquid_water": 1.0e3,
}


class _Component(GridComponent, StencilFactory):
    def __init__(self, grid, physical_constants=None, *, backend="b200", backend_options=None,
                 storage_shape=None, storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        self.grid = grid
        self.rpc = dict(DEFAULT_CONSTANTS)
        self.rpc.update(physical_constants or {})
        self.storage_shape = tuple(storage_shape or (grid.nx + 1, grid.ny + 1, grid.nz + 1))

    # what tasmania_b200.coupling needs to know about a component: its call convention (the
    # sympl base class in the reference) and the keys of its tendency_properties /
    # diagnostic_properties
    kind = "tendency"  # "tendency" | "implicit" (takes the timestep) | "diagnostic"
    tendency_names: tuple = ()
    diagnostic_names: tuple = ()

    def diagnostic_shape(self, name):
        return self.storage_shape

    default_physical_constants = DEFAULT_CONSTANTS

    @property
    def raw_physical_constants(self):
        """framework/base_components.py:L46-L48."""
        return dict(self.rpc)

    @property
    def _box(self):
        g = self.grid
        return dict(origin=(0, 0, 0), domain=(g.nx, g.ny, g.nz))


class KesslerMicrophysics(_Component):
    """Tendencies of qv, qc, qr and theta by autoconversion, accretion and rain evaporation."""

    def __init__(self, grid, air_pressure_on_interface_levels=True,
                 tendency_of_air_potential_temperature_in_diagnostics=False, rain_evaporation=True,
                 autoconversion_threshold=0.001, autoconversion_rate=0.001, collection_rate=2.2,
                 physical_constants=None, **kwargs):
        super().__init__(grid, physical_constants, **kwargs)
        self._pttd = tendency_of_air_potential_temperature_in_diagnostics
        self._apoil = air_pressure_on_interface_levels
        self._rain_evaporation = rain_evaporation
        self._a, self._k1, self._k2 = autoconversion_threshold, autoconversion_rate, collection_rate
        rd, rv = self.rpc["gas_constant_of_dry_air"], self.rpc["gas_constant_of_water_vapor"]
        self.backend_options.externals = {
            "air_pressure_on_interface_levels": air_pressure_on_interface_levels,
            "beta": rd / rv,
            "e": np.exp(1),
            "lhvw": self.rpc["latent_heat_of_vaporization_of_water"],
            "rain_evaporation": rain_evaporation,
            "set_output": self.get_subroutine_definition("set_output"),
        }
        self._stencil = self.compile_stencil("kessler")
        self._placeholder = self.zeros(shape=self.storage_shape)
        # kessler.py:L214-L253
        self.tendency_names = (mfcw, mfpw)
        if rain_evaporation:
            self.tendency_names += (mfwv,)
            if self._pttd:
                self.diagnostic_names = ("tendency_of_air_potential_temperature",)
            else:
                self.tendency_names += ("air_potential_temperature",)

    def array_call(self, state, out_tendencies, out_diagnostics, overwrite_tendencies):
        args = {
            "a": self._a, "k1": self._k1, "k2": self._k2,
            "in_rho": state["air_density"], "in_t": state["air_temperature"],
            "in_qc": state[mfcw], "in_qr": state[mfpw], "in_qv": state[mfwv],
            "out_qc_tnd": out_tendencies[mfcw], "out_qr_tnd": out_tendencies[mfpw],
            "out_qv_tnd": out_tendencies.get(mfwv, self._placeholder),
            "ow_out_qc_tnd": overwrite_tendencies[mfcw],
            "ow_out_qr_tnd": overwrite_tendencies[mfpw],
            "ow_out_qv_tnd": overwrite_tendencies.get(mfwv, False),
            "ow_out_theta_tnd": overwrite_tendencies.get("air_potential_temperature", False),
        }
        if self._apoil:
            args["in_p"] = state["air_pressure_on_interface_levels"]
            args["in_exn"] = state["exner_function_on_interface_levels"]
        else:
            args["in_p"], args["in_exn"] = state["air_pressure"], state["exner_function"]
        if self._rain_evaporation:
            args["out_theta_tnd"] = (out_diagnostics["tendency_of_air_potential_temperature"]
                                     if self._pttd else out_tendencies["air_potential_temperature"])
        else:
            args["out_theta_tnd"] = self._placeholder
        self._stencil(**args, **self._box)


class _Saturation(_Component):
    def __init__(self, grid, air_pressure_on_interface_levels=True, physical_constants=None, **kwargs):
        super().__init__(grid, physical_constants, **kwargs)
        self._apoil = air_pressure_on_interface_levels
        rd, rv = self.rpc["gas_constant_of_dry_air"], self.rpc["gas_constant_of_water_vapor"]
        self.backend_options.externals = {
            "air_pressure_on_interface_levels": air_pressure_on_interface_levels,
            "beta": rd / rv,
            "cp": self.rpc["specific_heat_of_dry_air_at_constant_pressure"],
            "e": np.exp(1),
            "lhvw": self.rpc["latent_heat_of_vaporization_of_water"],
            "rv": rv,
            "set_output": self.get_subroutine_definition("set_output"),
        }
        self._stencil = self.compile_stencil("saturation")

    def _p_exn(self, state):
        if self._apoil:
            return state["air_pressure_on_interface_levels"], state["exner_function_on_interface_levels"]
        return state["air_pressure"], state["exner_function"]


class KesslerSaturationAdjustmentDiagnostic(_Saturation):
    """Saturation adjustment as a diagnostic (implicit-tendency) component."""

    class_stencils = {"saturation": "saturation_diagnostic"}
    kind = "implicit"  # kessler.py:L588-L616
    tendency_names = ("air_potential_temperature",)
    diagnostic_names = (mfwv, mfcw, "air_temperature")

    def array_call(self, state, timestep, out_tendencies, out_diagnostics, overwrite_tendencies):
        in_p, in_exn = self._p_exn(state)
        self._stencil(
            in_p=in_p, in_t=state["air_temperature"], in_exn=in_exn, in_qv=state[mfwv],
            in_qc=state[mfcw], out_qv=out_diagnostics[mfwv], out_qc=out_diagnostics[mfcw],
            out_t=out_diagnostics["air_temperature"],
            tnd_theta=out_tendencies["air_potential_temperature"], dt=timestep.total_seconds(),
            ow_tnd_theta=overwrite_tendencies["air_potential_temperature"], **self._box)


class KesslerSaturationAdjustmentPrognostic(_Saturation):
    """Saturation adjustment as a tendency component with a saturation rate."""

    class_stencils = {"saturation": "saturation_prognostic"}
    tendency_names = (mfwv, mfcw, "air_potential_temperature")  # kessler.py:L917-L939

    def __init__(self, grid, air_pressure_on_interface_levels=True, saturation_rate=0.025,
                 physical_constants=None, **kwargs):
        super().__init__(grid, air_pressure_on_interface_levels, physical_constants, **kwargs)
        self._sr = saturation_rate

    def array_call(self, state, out_tendencies, out_diagnostics, overwrite_tendencies):
        in_p, in_exn = self._p_exn(state)
        self._stencil(
            in_p=in_p, in_t=state["air_temperature"], in_exn=in_exn, in_qv=state[mfwv],
            in_qc=state[mfcw], tnd_qv=out_tendencies[mfwv], tnd_qc=out_tendencies[mfcw],
            tnd_theta=out_tendencies["air_potential_temperature"], sr=self._sr,
            ow_tnd_qv=overwrite_tendencies[mfwv], ow_tnd_qc=overwrite_tendencies[mfcw],
            ow_tnd_theta=overwrite_tendencies["air_potential_temperature"], **self._box)


class KesslerFallVelocity(_Component):
    """Raindrop fall velocity."""

    kind = "diagnostic"
    diagnostic_names = ("raindrop_fall_velocity",)

    def __init__(self, grid, **kwargs):
        super().__init__(grid, **kwargs)
        self._in_rho_s = self.zeros(shape=self.storage_shape)
        self._stencil = self.compile_stencil("fall_velocity")

    def array_call(self, state, out):
        nx, ny, nz = self.grid.nx, self.grid.ny, self.grid.nz
        # slab broadcast by slice assignment, kessler.py:L1169
        self._in_rho_s[:nx, :ny, :nz] = state["air_density"][:nx, :ny, nz - 1:nz]
        self._stencil(in_rho=state["air_density"], in_rho_s=self._in_rho_s, in_qr=state[mfpw],
                      out_vt=out["raindrop_fall_velocity"], **self._box)


class KesslerSedimentation(_Component):
    """Tendency of qr due to sedimentation."""

    kind = "implicit"
    tendency_names = (mfpw,)

    def __init__(self, grid, sedimentation_flux_scheme="first_order_upwind", **kwargs):
        super().__init__(grid, **kwargs)
        if sedimentation_flux_scheme not in SEDIMENTATION_FLUX:
            raise ValueError(f"unknown sedimentation flux scheme {sedimentation_flux_scheme!r}")
        sflux = SEDIMENTATION_FLUX[sedimentation_flux_scheme]
        self.backend_options.externals = {
            "set_output": self.get_subroutine_definition("set_output"),
            "sflux": sflux,
            "sflux_extent": sflux.nb,
        }
        self._stencil = self.compile_stencil("sedimentation")

    def array_call(self, state, timestep, out_tendencies, out_diagnostics, overwrite_tendencies):
        self._stencil(in_rho=state["air_density"], in_h=state["height_on_interface_levels"],
                      in_qr=state[mfpw], in_vt=state["raindrop_fall_velocity"],
                      out_tnd_qr=out_tendencies[mfpw], ow_out_tnd_qr=overwrite_tendencies[mfpw],
                      **self._box)


class Clipping(_Component):
    """Negative values of the water species set to zero (physics/microphysics/utils.py:L58-L141): one
    ``clip`` launch per species over its whole storage, as the reference's ``array_call`` does."""

    kind = "diagnostic"

    def __init__(self, grid, water_species_names=None, **kwargs):
        super().__init__(grid, None, **kwargs)
        self._names = tuple(water_species_names or ())
        self.diagnostic_names = self._names
        self._stencil = self.compile_stencil("clip")

    def array_call(self, state, out):
        for name in self._names:
            out_q = out[name]
            self._stencil(in_field=state[name], out_field=out_q, origin=(0, 0, 0), domain=tuple(out_q.shape))


class Precipitation(_Component):
    """Precipitation rate and accumulated precipitation at the surface (2-D outputs)."""

    kind = "implicit"
    diagnostic_names = ("precipitation", "accumulated_precipitation")

    def diagnostic_shape(self, name):  # utils.py:L236-L247: one level
        return (self.storage_shape[0], self.storage_shape[1], 1)

    def __init__(self, grid, physical_constants=None, **kwargs):
        super().__init__(grid, physical_constants, **kwargs)
        self.backend_options.externals = {"rhow": self.rpc["density_of_liquid_water"]}
        self._stencil = self.compile_stencil("accumulated_precipitation")

    def array_call(self, state, timestep, out_tendencies, out_diagnostics, overwrite_tendencies):
        nx, ny, nz = self.grid.nx, self.grid.ny, self.grid.nz
        # sliced (non-contiguous) views of the surface level, utils.py:L270-L275
        self._stencil(
            in_rho=state["air_density"][:, :, nz - 1:nz], in_qr=state[mfpw][:, :, nz - 1:nz],
            in_vt=state["raindrop_fall_velocity"][:, :, nz - 1:nz],
            in_accprec=state["accumulated_precipitation"][:, :, :1],
            out_prec=out_diagnostics["precipitation"][:, :, :1],
            out_accprec=out_diagnostics["accumulated_precipitation"][:, :, :1],
            dt=timestep.total_seconds(), origin=(0, 0, 0), domain=(nx, ny, 1))
