# -*- coding: utf-8 -*-
"""Registration of the ``b200`` backend into tasmania's own registries.

    import tasmania            # the reference package, unmodified
    import tasmania_b200.plugin
    tasmania_b200.plugin.install()
    dycore = IsentropicDynamicalCore(domain, ..., backend="b200", ...)

After ``install()`` every ``StencilFactory`` of tasmania resolves ``backend="b200"`` like any
built-in backend:

* allocators ``zeros / ones / empty / as_storage`` (src/tasmania/framework/allocators.py:L40-L181)
  return ``B200Array`` device storages; ``as_storage("numpy", data=<B200Array>)`` (what
  ``to_numpy`` calls, framework/generic_functions.py:L35-L36) copies back to the host;
* ``stencil_compiler`` / ``subroutine_compiler`` (framework/stencil.py:L137-L204) get a b200
  entry: the stencil compiler snapshots ``backend_options.externals`` and swallows unused
  keyword arguments exactly like ``compiler_numpy`` + ``wrap``
  (framework/subclasses/stencil_compilers.py:L49-L99);
* the class-less stencils (``copy``, the ``math`` family, ``irelax``/``relax``, ``sts_rk*_0``,
  ``step_forward_euler[_momentum]``) are registered globally through
  ``StencilDefinition.register(handle, backend="b200", stencil=...)``
  (framework/stencil.py:L112-L130);
* the class-scoped ones (``diffusion``, ``smoothing``, ``damping``, ``montgomery``, ...; several
  classes define the same stencil name and only the per-instance registry tells them apart,
  framework/stencil.py:L460-L469) are attached to the reference classes as static methods
  tagged with ``tasmania.framework.tag.stencil_definition(backend="b200", stencil=...)``
  (framework/tag.py:L68-L80), before any instance is built;
* the flux / advection / sedimentation-flux *subroutines* only have to exist for the backend
  (``get_subroutine_definition``, framework/stencil.py:L379-L392): a fused CUDA kernel never
  calls them, so a descriptor function carrying the scheme is registered and travels to the
  b200 stencil through ``externals`` (e.g. rk3ws_si.py:L249-L264).

Nothing here computes: every definition marshals its arguments into one C-ABI call of
``libtasmania_b200.so`` (tasmania_b200/stencils.py).  No multi-backend dispatch, no CPU
fallback: a stencil without a b200 definition raises ``FactoryRegistryError`` as for any other
backend.
"""
from __future__ import annotations

import importlib
import inspect

import numpy as np

from tasmania_b200 import framework as fw
from tasmania_b200 import stencils as st
from tasmania_b200 import storage

BACKEND = fw.BACKEND
_installed = False


# ------------------------------------------------------------------ allocators
def _dtype(storage_options):
    return getattr(storage_options, "dtype", None) or np.float64


def _device(storage_options):
    return getattr(storage_options, "device", None) or storage.DEFAULT_DEVICE_OVERRIDE


def zeros_b200(shape, *, storage_options=None):
    return storage.zeros(shape, dtype=_dtype(storage_options), device=_device(storage_options))


def ones_b200(shape, *, storage_options=None):
    return storage.ones(shape, dtype=_dtype(storage_options), device=_device(storage_options))


def empty_b200(shape, *, storage_options=None):
    return storage.empty(shape, dtype=_dtype(storage_options), device=_device(storage_options))


def as_storage_b200(data, *, storage_options=None):
    return storage.as_storage(data, device=_device(storage_options))


subroutine_compiler_b200 = fw.subroutine_compiler_b200  # descriptors: nothing to compile


# ------------------------------------------------------------------ descriptors as functions
def _descriptor(scheme, name):
    """``_fill_registry`` only collects functions / methods (framework/stencil.py:L460-L469),
    so a scheme descriptor is wrapped in a function object that carries it."""

    def subroutine(*args, **kwargs):
        raise RuntimeError(
            f"'{name}' is a descriptor for the b200 backend: the scheme is evaluated inside the "
            "fused CUDA kernels, never as a separate subroutine")

    subroutine.__name__ = f"{name}_b200"
    subroutine.tb200_scheme = scheme
    return subroutine


def _order_bound(definition, key, value, name, **more):
    """A class-scoped b200 definition with the class's order (and, for the one-dimensional
    variants, its axis) baked into the externals."""

    def bound(externals, **kwargs):
        ext = dict(externals)
        ext[key] = value
        ext.update(more)
        return definition(ext, **kwargs)

    bound.__name__ = name
    bound.__signature__ = inspect.signature(definition)
    return bound


# (module, class, stencil name, b200 definition)
def _class_scoped_stencils():
    dv = "tasmania.isentropic.dynamics.diagnostics"
    dw = "tasmania.dwarfs.diagnostics"
    hd = "tasmania.dwarfs.subclasses.horizontal_diffusers"
    hs = "tasmania.dwarfs.subclasses.horizontal_smoothers"
    out = [
        (dv, "IsentropicDiagnostics", "diagnostic_variables", st.diagnostic_variables_b200),
        (dv, "IsentropicDiagnostics", "montgomery", st.montgomery_b200),
        (dv, "IsentropicDiagnostics", "height", st.height_b200),
        (dv, "IsentropicDiagnostics", "density_and_temperature", st.density_and_temperature_b200),
        (dw, "HorizontalVelocity", "momenta", st.momenta_b200),
        (dw, "HorizontalVelocity", "velocity_x", st.velocity_x_b200),
        (dw, "HorizontalVelocity", "velocity_y", st.velocity_y_b200),
        (dw, "WaterConstituent", "density", st.density_b200),
        (dw, "WaterConstituent", "mass_fraction", st.mass_fraction_b200),
    ]
    # the 2-D schemes and their ..._1dx / ..._1dy variants (same modules)
    for mod, cls, order in (("second_order", "SecondOrder", 2), ("fourth_order", "FourthOrder", 4)):
        for sfx, axis in (("", None), ("1DX", 0), ("1DY", 1)):
            out.append((f"{hd}.{mod}", cls + sfx, "diffusion",
                        _order_bound(st.diffusion_b200, "diffusion_order", order,
                                     f"diffusion_{mod}{'_' + sfx.lower() if sfx else ''}_b200", diffusion_axis=axis)))
    for mod, cls, order in (("first_order", "FirstOrder", 1), ("second_order", "SecondOrder", 2),
                            ("third_order", "ThirdOrder", 3)):
        for sfx, axis in (("", None), ("1DX", 0), ("1DY", 1)):
            out.append((f"{hs}.{mod}", cls + sfx, "smoothing",
                        _order_bound(st.smoothing_b200, "smoothing_order", order,
                                     f"smoothing_{mod}{'_' + sfx.lower() if sfx else ''}_b200", smoothing_axis=axis)))
    out += [
        ("tasmania.dwarfs.subclasses.vertical_dampers.rayleigh", "Rayleigh", "damping",
         st.damping_b200),
        ("tasmania.burgers.dynamics.stepper", "BurgersStepper", "forward_euler",
         st.burgers_forward_euler_b200),
        ("tasmania.isentropic.physics.vertical_advection", "IsentropicVerticalAdvection", "stencil",
         st.vertical_advection_b200),
        ("tasmania.isentropic.physics.coriolis", "IsentropicConservativeCoriolis", "coriolis",
         st.coriolis_b200),
        ("tasmania.isentropic.physics.implicit_vertical_advection",
         "IsentropicImplicitVerticalAdvectionDiagnostic", "implicit_vertical_advection",
         st.implicit_vertical_advection_b200),
        ("tasmania.isentropic.physics.implicit_vertical_advection",
         "IsentropicImplicitVerticalAdvectionPrognostic", "stencil",
         st.implicit_vertical_advection_tendency_b200),
        ("tasmania.physics.turbulence", "Smagorinsky2d", "smagorinsky", st.smagorinsky_b200),
        ("tasmania.isentropic.physics.turbulence", "IsentropicSmagorinsky", "smagorinsky",
         st.smagorinsky_isentropic_b200),
    ]
    out += [(m, c, s, getattr(st, d)) for (m, c, s, d) in st.KESSLER_CLASS_STENCILS]
    return out


def _class_scoped_subroutines():
    mf = "tasmania.isentropic.dynamics.subclasses.minimal_horizontal_fluxes"
    ff = "tasmania.isentropic.dynamics.subclasses.horizontal_fluxes"
    vf = "tasmania.isentropic.dynamics.subclasses.minimal_vertical_fluxes"
    ad = "tasmania.burgers.dynamics.subclasses.advection"
    out = []
    for mod, cls, scheme in (("upwind", "Upwind", "upwind"), ("centered", "Centered", "centered"),
                             ("third_order_upwind", "ThirdOrderUpwind", "third_order_upwind"),
                             ("fifth_order_upwind", "FifthOrderUpwind", "fifth_order_upwind")):
        for pkg in (mf, ff, vf):
            for name in ("flux_dry", "flux_moist"):
                out.append((f"{pkg}.{mod}", cls, name, _descriptor(st.FLUX[scheme], name)))
    for mod, cls in (("first_order", "FirstOrder"), ("second_order", "SecondOrder"),
                     ("third_order", "ThirdOrder"), ("fourth_order", "FourthOrder"),
                     ("fifth_order", "FifthOrder"), ("sixth_order", "SixthOrder")):
        out.append((f"{ad}.{mod}", cls, "advection", _descriptor(st.ADVECTION[mod], "advection")))
    sf = "tasmania.physics.microphysics.sedimentation_fluxes"
    for mod, cls, name in (("first_order", "FirstOrderUpwind", "first_order_upwind"),
                           ("second_order", "SecondOrderUpwind", "second_order_upwind")):
        out.append((f"{sf}.{mod}", cls, "flux", _descriptor(st.SEDIMENTATION_FLUX[name], "flux")))
    out.append(("tasmania.physics.turbulence", "Smagorinsky2d", "smagorinsky_core",
                _descriptor("smagorinsky_core", "smagorinsky_core")))
    return out


GLOBAL_STENCILS = (
    "copy", "copychange", "abs", "iabs", "add", "iadd", "addsub", "iaddsub", "clip", "iclip", "fma",
    "mul", "imul", "scale", "iscale", "sub", "isub", "sts_rk2_0", "sts_rk3ws_0", "irelax", "relax",
    "step_forward_euler", "step_forward_euler_momentum", "thomas",
)


def _batch_dict_operator_fma():
    """Give the reference's ``DataArrayDictOperator.fma`` (src/tasmania/utils/xarrayx.py:L688-L740)
    the one-launch-per-stage form on backend b200: same dictionary / units logic, but the per-field
    ``fma`` stencil calls are collected and issued as one ``tb200_fma_fields`` launch.  Every other
    backend goes through the original method.  Returns what was replaced (for the report)."""
    from tasmania.utils import xarrayx
    from tasmania.utils.storage import deepcopy_dataarray

    original = xarrayx.DataArrayDictOperator.fma
    if getattr(original, "__tasmania_b200__", False):
        return None

    def fma(self, dict1, dict2, factor, out=None, field_properties=None):
        if getattr(self, "backend", None) != BACKEND:
            return original(self, dict1, dict2, factor, out=out, field_properties=field_properties)
        field_properties = field_properties or {}
        out = out or {}
        if "time" in dict1 or "time" in dict2:
            out["time"] = dict1.get("time", dict2.get("time", None))
        shared_keys = set(dict1.keys()).intersection(dict2.keys()).difference(("time",))
        outs, ins_a, ins_b = [], [], []
        for key in shared_keys:
            props = field_properties.get(key, {})
            units = props.get("units", dict1[key].attrs["units"])
            field1 = dict1[key].to_units(units)
            rfield2 = dict2[key].to_units(units).data
            if key in out:
                out[key].attrs["units"] = units
            else:
                out[key] = deepcopy_dataarray(field1)
            outs.append(out[key].data)
            ins_a.append(field1.data)
            ins_b.append(rfield2)
        by_shape = {}
        for o, a, b in zip(outs, ins_a, ins_b):
            by_shape.setdefault(tuple(o.shape), []).append((o, a, b))
        for shape, group in by_shape.items():
            st.fma_fields([g[0] for g in group], [g[1] for g in group], [g[2] for g in group], factor,
                           origin=(0, 0, 0), domain=shape)
        return out

    fma.__tasmania_b200__ = True
    xarrayx.DataArrayDictOperator.fma = fma
    return "DataArrayDictOperator.fma"


def _frame_relaxed_enforce_raw():
    """Give the reference's ``Relaxed`` boundary the one-launch form of ``enforce_raw``
    (src/tasmania/domain/horizontal_boundary.py:L299-L344 over relaxed.py:L119-L160) on backend
    b200: the same selection of fields, units and extents, relaxed by ``tb200_relax_frame`` on the
    frame where the object's own coefficient matrix is non-zero instead of one full-box ``irelax``
    per field.  Other backends go through the original method."""
    from tasmania.domain.subclasses.horizontal_boundaries.relaxed import Relaxed
    from tasmania_b200 import boundary as b200_boundary

    original = Relaxed.enforce_raw
    if getattr(original, "__tasmania_b200__", False):
        return None

    def enforce_raw(self, state, field_properties=None):
        if getattr(self, "backend", None) != BACKEND:
            return original(self, state, field_properties)
        rfps = {name: {"units": self.reference_state[name].attrs["units"]}
                for name in self.reference_state if name != "time"}
        fps = rfps if field_properties is None else {
            key: val for key, val in field_properties.items() if key in rfps}
        names = [name for name in state if name != "time" and name in fps]
        if not names:
            return
        box = getattr(self, "_b200_free_box", None)
        if box is None:  # once per object: where its own gamma vanishes
            box = b200_boundary.Relaxed._gamma_free_box(storage.to_numpy(self._gamma)[:, :, 0])
            self._b200_free_box = box
        refs, extents = [], []
        for name in names:
            units = fps[name].get("units", rfps[name]["units"])
            refs.append(self.reference_state[name].to_units(units).data)
            extents.append((
                self.ni + 1 if "at_u_locations" in name or "at_uv_locations" in name else self.ni,
                self.nj + 1 if "at_v_locations" in name or "at_uv_locations" in name else self.nj,
                self.physical_grid.nz + 1 if "on_interface_levels" in name else self.physical_grid.nz))
        b200_boundary.relax_frame([state[n] for n in names], refs, extents, self._gamma, box)

    enforce_raw.__tasmania_b200__ = True
    Relaxed.enforce_raw = enforce_raw
    return "Relaxed.enforce_raw"


def _fused_stage(moist=False):
    """Put the fused RK stage (``tb200_isentropic_stage_dry``: three kernels instead of the 13
    stencil launches of a stage) behind the reference's own
    ``IsentropicDynamicalCore.stage_array_call_dry`` (src/tasmania/isentropic/dynamics/
    dycore.py:L641-L721) on backend b200.  The patched method keeps the reference's bookkeeping
    (current-solution pointers of the dycore and of its prognostic object, topography upload,
    time label) and hands everything else to the library; it falls back to the original method
    -- the per-stencil b200 kernels -- whenever the configuration is outside the fused kernels'
    scope: fast tendencies, slow tendencies other than those of s, su, sv (dry stage), a boundary
    other than the 2-D ``Relaxed`` or the 2-D ``Periodic``, a prognostic scheme
    other than RK3WSSI / ForwardEulerSI, reference fields in non-canonical units, foreign storages.

    ``moist=True`` does the same for ``stage_array_call_moist`` (dycore.py:L723-L843) with
    ``tb200_isentropic_stage_moist``: the dry kernels plus one kernel for the three water
    constituents (density, K1 share, mass fraction, lateral relaxation), no sq* storages.

    Intermediate stages neither write nor read u, v (``skip_uv_out`` / ``derive_uv_in``: their
    successor re-diagnoses them with the formula ``get_velocity_components`` ends every stage with,
    so the step's result is bit-identical) unless the dycore has a fast tendency or diagnostic
    component, which may look at an intermediate state's velocities."""
    from tasmania.isentropic.dynamics.dycore import IsentropicDynamicalCore
    from tasmania_b200 import lib

    method = "stage_array_call_moist" if moist else "stage_array_call_dry"
    original = getattr(IsentropicDynamicalCore, method)
    if getattr(original, "__tasmania_b200__", False):
        return None

    S, SU, SV = "air_isentropic_density", "x_momentum_isentropic", "y_momentum_isentropic"
    U, V, MTG = "x_velocity_at_u_locations", "y_velocity_at_v_locations", "montgomery_potential"
    QN = ("mass_fraction_of_water_vapor_in_air", "mass_fraction_of_cloud_liquid_water_in_air",
          "mass_fraction_of_precipitation_water_in_air")
    UNITS = {S: "kg m^-2 K^-1", SU: "kg m^-1 K^-1 s^-1", SV: "kg m^-1 K^-1 s^-1", U: "m s^-1", V: "m s^-1"}
    if moist:
        UNITS.update({q: "g g^-1" for q in QN})
    SUBSTEPS = {  # stage -> (increment of the time label, stage time step), the reference's own
        # timedelta arithmetic (rk3ws_si.py:L115-L123, forward_euler_si.py:L106-L111)
        "rk3ws_si": (lambda t: (t / 3.0, t / 3.0), lambda t: (t / 6.0, 0.5 * t), lambda t: (0.5 * t, t)),
        "forward_euler_si": (lambda t: (t, t),),
    }

    def plan(self, tendencies):
        """The static part of the decision, cached on the object; None = not fusable."""
        cached = getattr(self, "_b200_fused_plan", None)
        if cached is not None:
            return cached or None
        ok = getattr(self, "backend", None) == BACKEND and bool(getattr(self, "_moist", not moist)) == moist
        hb, pr = self.horizontal_boundary, self._prognostic
        periodic = type(hb).__name__ == "Periodic"  # the stage wraps s itself (cfg.periodic)
        ok = ok and (type(hb).__name__ == "Relaxed" or periodic) and getattr(type(pr), "name", None) in SUBSTEPS
        scheme = getattr(getattr(pr, "_hflux", None), "name", None)
        ok = ok and scheme in lib.FLUX_SCHEMES and hasattr(pr, "_diagnostics")
        if ok:
            ref = hb.reference_state
            ok = all(n in ref and ref[n].attrs.get("units") == UNITS[n] for n in UNITS)
        if not ok:
            self._b200_fused_plan = False
            return None
        diag = pr._diagnostics
        g = self.grid
        rpc = diag.rpc
        shape = tuple(diag._topo.shape)
        self._b200_fused_plan = {
            "scheme": lib.FLUX_SCHEMES[scheme], "substeps": SUBSTEPS[type(pr).name],
            "dx": g.dx.to_units("m").values.item(), "dy": g.dy.to_units("m").values.item(),
            "dz": g.dz.to_units("K").values.item(),
            "theta_s": float(g.z_on_interface_levels.to_units("K").values[-1]),
            "constants": [rpc["air_pressure_at_sea_level"], rpc["gas_constant_of_dry_air"],
                          rpc["gravitational_acceleration"],
                          rpc["specific_heat_of_dry_air_at_constant_pressure"]],
            # the stage's hand-off arrays: a library context held by the dycore object (tb200_ctx)
            "scratch_owner": None, "scratch": None,
            "lazy": (self.fast_tendency_component is None and self.fast_diagnostic_component is None
                     and bool(lib.load().tb200_stage_lazy_velocities(int(g.nz)))),
            "periodic": periodic,
        }
        self._b200_fused_plan["scratch_owner"], self._b200_fused_plan["scratch"] = storage.stage_scratch(
            shape, 3, _device(self.storage_options))
        if periodic:
            if not lib.load().tb200_stage_lazy_velocities(int(g.nz)):  # an earlier kernel variant forced by the environment
                self._b200_fused_plan = False
                return None
            self._b200_fused_plan["gamma"] = storage.zeros(shape[:2] + (1,), device=_device(self.storage_options))
        return self._b200_fused_plan

    def stage_array_call(self, stage, state, tendencies, timestep, out_state):
        # slow tendencies of s, su, sv ride the fused DRY stage (tb200_isentropic_stage.s_tnd ...);
        # anything else (moist stage, tendencies of the water constituents, foreign storages, an
        # earlier kernel variant forced by the environment) takes the per-stencil kernels
        slow = {k: v for k, v in (tendencies or {}).items() if k != "time"}
        if slow and (moist or not set(slow) <= {S, SU, SV}
                     or not all(isinstance(v, storage.B200Array) and tuple(v.shape) == tuple(state[S].shape)
                                for v in slow.values())
                     or not lib.load().tb200_stage_lazy_velocities(int(self.grid.nz))):
            return original(self, stage, state, tendencies, timestep, out_state)
        p = plan(self, tendencies)
        needed = (S, SU, SV, U, V, MTG) + (QN if moist else ())
        if p is None or not all(isinstance(state[n], storage.B200Array) for n in needed):
            return original(self, stage, state, tendencies, timestep, out_state)
        hb, pr = self.horizontal_boundary, self._prognostic
        diag = pr._diagnostics
        g = self.grid
        nx, ny, nz = g.nx, g.ny, g.nz
        ref = hb.reference_state
        if stage == 0:  # the reference's bookkeeping, dycore.py:L662-L668 and rk3ws_si.py:L126-L130
            self._s_now, self._su_now, self._sv_now = state[S], state[SU], state[SV]
            pr._s_now, pr._mtg_now, pr._su_now, pr._sv_now = state[S], state[MTG], state[SU], state[SV]
            if moist:  # the step-start mass fractions: the kernel forms s q itself (dycore.py:L762-L779)
                self._b200_q_now = [state[q] for q in QN]
        # the topography of the moment, as get_montgomery_potential uploads it (diagnostics.py:L216-L219)
        diag._topo[:nx, :ny, nz] = diag.as_storage(data=g.topography.profile.to_units("m").values)
        dtr, dt = p["substeps"][stage](timestep)
        damp = bool(self._damp and (self._damp_at_every_stage or stage == self.stages - 1))
        cfg = lib.StageCfg()
        cfg.nx, cfg.ny, cfg.nz, cfg.nb = nx, ny, nz, hb.nb
        cfg.flux_scheme = p["scheme"]
        cfg.damp = int(damp)
        cfg.dt, cfg.dt_full = dt.total_seconds(), timestep.total_seconds()
        cfg.dx, cfg.dy, cfg.dz, cfg.eps = p["dx"], p["dy"], p["dz"], pr._eps
        cfg.pt, cfg.theta_s = pr._pt, p["theta_s"]
        cfg.constants[:] = p["constants"]
        cfg.part = 0
        cfg.rim[:] = [0, 0, 0, 0]
        cfg.derive_uv_in = int(p["lazy"] and stage > 0)
        cfg.skip_uv_out = int(p["lazy"])  # the last stage's u, v: one tb200_velocity_components pass below
        periodic = p["periodic"]
        if periodic:  # the wrap of s, su, sv comes between the momentum step and the damping (dycore.py:L684-L700)
            cfg.periodic, cfg.damp, cfg.skip_uv_out = 1, 0, 1
        scratch_s = out_state[S] if cfg.skip_uv_out else p["scratch"][2]
        f = lib.as_field
        keep_tnd = None
        if slow:  # a missing one is a field of zeros: x - 0.0 == x, the reference's own form
            import ctypes as C

            if "zero_tnd" not in p:
                p["zero_tnd"] = storage.zeros(tuple(state[S].shape), device=_device(self.storage_options))
            keep_tnd = [f(slow.get(n, p["zero_tnd"])) for n in (S, SU, SV)]
            cfg.s_tnd, cfg.su_tnd, cfg.sv_tnd = (C.pointer(k) for k in keep_tnd)
        rmat = self._damper._rmat if self._damp else None
        args = (cfg, f(pr._s_now), f(pr._su_now), f(pr._sv_now), f(pr._mtg_now),
                f(state[S]), f(state[SU]), f(state[SV]), f(state[U]), f(state[V]),
                f(out_state[S]), f(out_state[SU]), f(out_state[SV]), f(out_state[U]), f(out_state[V]),
                f(ref[S].data), f(ref[SU].data), f(ref[SV].data), f(ref[U].data), f(ref[V].data),
                f(p["gamma"] if periodic else hb._gamma), f(rmat), f(diag._topo[:, :, nz:nz + 1]),
                f(p["scratch"][0]), f(p["scratch"][1]), f(scratch_s))
        if moist:
            import ctypes as C

            keep = [[f(x) for x in xs] for xs in (self._b200_q_now, [state[q] for q in QN],
                                                   [out_state[q] for q in QN], [ref[q].data for q in QN])]
            arrays = [(lib.FieldP * 3)(*[C.pointer(k) for k in ks]) for ks in keep]
            rc = lib.load().tb200_isentropic_stage_moist(*args, *arrays, lib.current_stream())
            lib.check(rc, "tb200_isentropic_stage_moist")
            del keep
        else:
            rc = lib.load().tb200_isentropic_stage_dry(*args, lib.current_stream())
            lib.check(rc, "tb200_isentropic_stage_dry")
        if periodic:
            for n in (S, SU, SV) + (QN if moist else ()):  # hb.enforce_raw: periodic.py:L98-L122
                lib.check(lib.load().tb200_periodic_enforce(f(out_state[n]), hb.nx, hb.ny, hb.nb, hb.nx, hb.ny,
                                                            lib.current_stream()), "tb200_periodic_enforce")
            if damp:  # dycore.py:L694-L700
                self._damper(timestep, self._s_now, out_state[S], ref[S].data, out_state[S])
                self._damper(timestep, self._su_now, out_state[SU], ref[SU].data, out_state[SU])
                self._damper(timestep, self._sv_now, out_state[SV], ref[SV].data, out_state[SV])
            if not p["lazy"] or stage == self.stages - 1:  # dycore.py:L702-L721
                self._velocity_components.get_velocity_components(
                    out_state[S], out_state[SU], out_state[SV], out_state[U], out_state[V])
                hb.set_outermost_layers_x(out_state[U], field_name=U)
                hb.set_outermost_layers_y(out_state[V], field_name=V)
        elif p["lazy"] and stage == self.stages - 1:
            rc = lib.load().tb200_velocity_components(
                f(out_state[S]), f(out_state[SU]), f(out_state[SV]), f(out_state[U]), f(out_state[V]),
                f(ref[U].data), f(ref[V].data), nx, ny, nz, lib.current_stream())
            lib.check(rc, "tb200_velocity_components")
        out_state["time"] = state["time"] + dtr

    stage_array_call.__name__ = method
    stage_array_call.__tasmania_b200__ = True
    stage_array_call.__wrapped_original__ = original
    setattr(IsentropicDynamicalCore, method, stage_array_call)
    return "IsentropicDynamicalCore." + method


def _fused_stage_dry():
    return _fused_stage(False)


def _fused_stage_moist():
    return _fused_stage(True)



def install(strict: bool = False, batch_fma: bool = True, frame_relax: bool = True,
            fused_stage: bool = True) -> dict:
    """Register the b200 backend into the importable ``tasmania`` package.  Returns a report
    ``{"global": [...], "class_scoped": [...], "skipped": [...]}``; with ``strict`` a reference
    class that cannot be imported raises instead of being skipped (sub-packages pulling
    optional dependencies may be unavailable)."""
    global _installed
    from tasmania.framework import allocators as ta
    from tasmania.framework import stencil as ts
    from tasmania.framework import tag as tt

    report = {"global": [], "class_scoped": [], "skipped": []}
    if _installed:
        return report

    # 1. allocators
    ta.zeros.register(zeros_b200, backend=BACKEND)
    ta.ones.register(ones_b200, backend=BACKEND)
    ta.empty.register(empty_b200, backend=BACKEND)
    ta.as_storage.register(as_storage_b200, backend=BACKEND)
    try:  # to_numpy(<B200Array>): an overload of the reference's singledispatch converter
        from tasmania.framework.subclasses.allocators.as_storage_numpy import as_storage_numpy

        @as_storage_numpy.register
        def _(data: storage.B200Array, *, storage_options=None):
            return data.to_numpy()
    except Exception as exc:  # pragma: no cover - depends on optional deps of the reference
        if strict:
            raise
        report["skipped"].append(("as_storage_numpy", repr(exc)))

    # 2. compilers
    ts.StencilCompiler.register(fw.compiler_b200, backend=BACKEND)
    ts.SubroutineCompiler.register(subroutine_compiler_b200, backend=BACKEND)

    # 3. class-less stencils and subroutines
    for name in GLOBAL_STENCILS:
        ts.StencilDefinition.register(fw.get_stencil_definition(name), backend=BACKEND, stencil=name)
        report["global"].append(name)
    # the reference's class-less `diffusion` (a hyperdiffusion filter) is `hyperdiffusion` in the
    # mirrors' flat registry; class-scoped `diffusion` definitions (the dwarfs) take precedence on
    # their instances exactly as they do for the numpy backend
    ts.StencilDefinition.register(fw.get_stencil_definition("hyperdiffusion"), backend=BACKEND,
                                  stencil="diffusion")
    report["global"].append("diffusion")
    for name in ("set_output", "thomas", "setup_thomas", "setup_thomas_bc"):
        ts.SubroutineDefinition.register(_descriptor(name, name), backend=BACKEND, stencil=name)

    # 4. class-scoped definitions: tagged static methods on the reference classes
    def attach(modname, clsname, stencil, fn, tagger):
        try:
            cls = getattr(importlib.import_module(modname), clsname)
        except Exception as exc:
            if strict:
                raise
            report["skipped"].append((f"{modname}.{clsname}:{stencil}", repr(exc)))
            return
        tagged = tagger(backend=BACKEND, stencil=stencil)(fn)
        setattr(cls, f"_{stencil}_b200", staticmethod(tagged))
        report["class_scoped"].append(f"{clsname}:{stencil}")

    for modname, clsname, stencil, fn in _class_scoped_stencils():
        attach(modname, clsname, stencil, fn, tt.stencil_definition)
    for modname, clsname, stencil, fn in _class_scoped_subroutines():
        attach(modname, clsname, stencil, fn, tt.subroutine_definition)

    # 5. the stage update of the reference's tendency steppers in one launch (bit-identical)
    if batch_fma:
        try:
            patched = _batch_dict_operator_fma()
            if patched:
                report["batched"] = [patched]
        except Exception as exc:  # pragma: no cover - depends on optional deps of the reference
            if strict:
                raise
            report["skipped"].append(("DataArrayDictOperator.fma", repr(exc)))

    if frame_relax:
        try:
            patched = _frame_relaxed_enforce_raw()
            if patched:
                report.setdefault("batched", []).append(patched)
        except Exception as exc:  # pragma: no cover - depends on optional deps of the reference
            if strict:
                raise
            report["skipped"].append(("Relaxed.enforce_raw", repr(exc)))

    # 6. the fused RK stage behind the reference's own dynamical core
    if fused_stage:
        try:
            for patched in (_fused_stage_dry(), _fused_stage_moist()):
                if patched:
                    report.setdefault("fused", []).append(patched)
        except Exception as exc:  # pragma: no cover - depends on optional deps of the reference
            if strict:
                raise
            report["skipped"].append(("IsentropicDynamicalCore.stage_array_call_dry / _moist", repr(exc)))

    _installed = True
    return report
