# -*- coding: utf-8 -*-
"""Physics-dynamics coupling on b200 storages ("coupler glue on device"):
host-side mirror, at the raw-array level, of

  TimeIntegrationOptions            src/tasmania/framework/options.py
  TendencyStepper (forward_euler,   src/tasmania/framework/steppers.py:L40-L140,
    rk2, rk3ws)                     src/tasmania/framework/subclasses/tendency_steppers/*.py
  ConcurrentCoupling (serial)       src/tasmania/framework/concurrent_coupling.py:L246-L374,
                                    concurrent_coupling_utils.py:L72-L83
  SequentialUpdateSplitting         src/tasmania/framework/sequential_update_splitting.py:L97-L194
  FromDiagnosticToTendency /        src/tasmania/framework/promoter.py:L161-L176, L290-L305
    FromTendencyToDiagnostic
  AirPotentialTemperatureTo*        src/tasmania/isentropic/utils.py:L27-L62
  IsentropicDiagnostics,            src/tasmania/isentropic/physics/diagnostics.py:L41-L301
    IsentropicVelocityComponents      (the sympl components around the cores of isentropic.py / dwarfs.py)
  IsentropicHorizontalSmoothing     src/tasmania/isentropic/physics/horizontal_smoothing.py:L41-L180
  IsentropicHorizontalDiffusion     src/tasmania/isentropic/physics/horizontal_diffusion.py:L41-L253

with the same class names and call signatures, minus the DataArray / units layer (every
isentropic component already works in the same units, so ``DataArrayDictOperator``'s conversions
are identities on this path).  A state is a dict name -> B200Array (+ ``"time"``).

What is b200-specific: every stage of a tendency stepper is ONE kernel launch over all stepped
fields (``tb200_fma_fields``) instead of one ``fma`` stencil call per field, buffers are
allocated once and swapped like ``DataArrayDictOperator.update_swap`` does, and nothing here
touches the host: the whole physics suite of the moist benchmark runs as a stream of kernels.
"""
from __future__ import annotations

import os

from tasmania_b200 import stencils
from tasmania_b200.dwarfs import HorizontalDiffusion, HorizontalSmoothing, HorizontalVelocity
from tasmania_b200.framework import BackendOptions, GridComponent, StencilFactory, StorageOptions
from tasmania_b200.isentropic import MTG, S, SU, SV, U, V
from tasmania_b200.isentropic import IsentropicDiagnostics as _DiagnosticsCore

mfwv = "mass_fraction_of_water_vapor_in_air"
mfcw = "mass_fraction_of_cloud_liquid_water_in_air"
mfpw = "mass_fraction_of_precipitation_water_in_air"
P = "air_pressure_on_interface_levels"
EXN = "exner_function_on_interface_levels"
H = "height_on_interface_levels"


def _state_shape(state):
    """Shape of the 3-D storages of a state (the largest one: 2-D diagnostics have one level)."""
    best = None
    for name, arr in state.items():
        if name != "time" and (best is None or arr.shape[2] > best[2]):
            best = tuple(arr.shape)
    return best


def update_swap(dst, src):
    """DataArrayDictOperator.update_swap (src/tasmania/utils/xarrayx.py): the arrays of ``src``
    go into ``dst``; the ones they replace go back into ``src`` to be reused as output buffers."""
    for name in list(src):
        if name == "time":
            continue
        old = dst.get(name)
        dst[name] = src[name]
        if old is not None:
            src[name] = old
        else:
            del src[name]


# ------------------------------------------------------------------------------ components
class _DomainComponent(GridComponent, StencilFactory):
    kind = "diagnostic"
    tendency_names: tuple = ()
    diagnostic_names: tuple = ()

    def __init__(self, grid, *, backend="b200", backend_options=None, storage_shape=None,
                 storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        self.grid = grid
        self.storage_shape = tuple(storage_shape or (grid.nx + 1, grid.ny + 1, grid.nz + 1))

    def diagnostic_shape(self, name):
        return self.storage_shape


class IsentropicDiagnostics(_DomainComponent):
    """p, exn, mtg, h (and rho, T when ``moist``) from the isentropic density."""

    def __init__(self, grid, moist, pt, **kwargs):
        super().__init__(grid, **kwargs)
        self._moist, self._pt = moist, float(pt)
        self._core = _DiagnosticsCore(grid, backend_options=BackendOptions(),
                                      storage_shape=self.storage_shape,
                                      storage_options=self.storage_options)
        self.diagnostic_names = (P, EXN, H, MTG) + (("air_density", "air_temperature") if moist else ())

    def array_call(self, state, out):
        self._core.get_diagnostic_variables(state[S], self._pt, out[P], out[EXN], out[MTG], out[H])
        if self._moist:
            self._core.get_density_and_temperature(state[S], out[EXN], out[H], out["air_density"],
                                                   out["air_temperature"])


class IsentropicVelocityComponents(_DomainComponent):
    """u, v from s, su, sv, outermost faces from the lateral boundary."""

    diagnostic_names = (U, V)

    def __init__(self, grid, horizontal_boundary, **kwargs):
        super().__init__(grid, **kwargs)
        self.horizontal_boundary = horizontal_boundary
        self._core = HorizontalVelocity(grid, staggering=True, backend_options=BackendOptions(),
                                        storage_options=self.storage_options)

    def array_call(self, state, out):
        hb = self.horizontal_boundary
        self._core.get_velocity_components(state[S], state[SU], state[SV], out[U], out[V])
        hb.set_outermost_layers_x(out[U], field_name=U, time=state.get("time"))
        hb.set_outermost_layers_y(out[V], field_name=V, time=state.get("time"))


class IsentropicHorizontalSmoothing(_DomainComponent):
    """Horizontal smoothing of s, su, sv (and of the water species)."""

    def __init__(self, grid, nb, smooth_type, smooth_coeff, smooth_coeff_max, smooth_damp_depth,
                 moist=False, smooth_moist_coeff=None, smooth_moist_coeff_max=None,
                 smooth_moist_damp_depth=None, **kwargs):
        super().__init__(grid, **kwargs)
        self._moist = moist and smooth_moist_coeff is not None
        make = lambda c, cmax, depth: HorizontalSmoothing.factory(  # noqa: E731
            smooth_type, self.storage_shape, c, cmax, depth, nb, backend_options=BackendOptions(),
            storage_options=self.storage_options)
        self._core = make(smooth_coeff, smooth_coeff_max, smooth_damp_depth)
        self.diagnostic_names = (S, SU, SV)
        if self._moist:
            cmax = smooth_moist_coeff if smooth_moist_coeff_max is None else smooth_moist_coeff_max
            self._core_moist = make(smooth_moist_coeff, cmax, smooth_moist_damp_depth or 0)
            self.diagnostic_names += (mfwv, mfcw, mfpw)

    def array_call(self, state, out):
        for n in (S, SU, SV):
            self._core(state[n], out[n])
        if self._moist:
            for n in (mfwv, mfcw, mfpw):
                self._core_moist(state[n], out[n])


class IsentropicHorizontalDiffusion(_DomainComponent):
    """Horizontal numerical diffusion of s, su, sv (and of the water species) as a tendency
    component: src/tasmania/isentropic/physics/horizontal_diffusion.py:L41-L253 around the
    HorizontalDiffusion dwarf (row K8 of SURVEY.md section 8a)."""

    kind = "tendency"

    def __init__(self, grid, nb, diffusion_type, diffusion_coeff, diffusion_coeff_max, diffusion_damp_depth,
                 moist=False, diffusion_moist_coeff=None, diffusion_moist_coeff_max=None,
                 diffusion_moist_damp_depth=None, **kwargs):
        super().__init__(grid, **kwargs)
        self._moist = moist and diffusion_moist_coeff is not None
        make = lambda c, cmax, depth: HorizontalDiffusion.factory(  # noqa: E731
            diffusion_type, self.storage_shape, grid.dx, grid.dy, c, cmax, depth, nb,
            backend_options=BackendOptions(), storage_options=self.storage_options)
        self._core = make(diffusion_coeff, diffusion_coeff_max, diffusion_damp_depth)
        self.tendency_names = (S, SU, SV)
        if self._moist:
            cmax = diffusion_moist_coeff if diffusion_moist_coeff_max is None else diffusion_moist_coeff_max
            self._core_moist = make(diffusion_moist_coeff, cmax, diffusion_moist_damp_depth or 0)
            self.tendency_names += (mfwv, mfcw, mfpw)

    def array_call(self, state, out_tendencies, out_diagnostics=None, overwrite_tendencies=None):
        ow = overwrite_tendencies or {}
        for n in (S, SU, SV):
            self._core(state[n], out_tendencies[n], overwrite_output=ow.get(n, True))
        if self._moist:
            for n in (mfwv, mfcw, mfpw):
                self._core_moist(state[n], out_tendencies[n], overwrite_output=ow.get(n, True))


class FromDiagnosticToTendency(_DomainComponent):
    """``diagnostic_name`` of the state copied (on the grid box) into the tendency ``tendency_name``."""

    kind = "d2t"

    def __init__(self, grid, diagnostic_name, tendency_name=None, **kwargs):
        super().__init__(grid, **kwargs)
        self.diagnostic_name = diagnostic_name
        self.tendency_name = tendency_name or diagnostic_name.replace("tendency_of_", "")
        self.tendency_names = (self.tendency_name,)
        self._stencil_copy = self.compile_stencil("copy")

    def array_call(self, diagnostics, out):
        g = self.grid
        self._stencil_copy(src=diagnostics[self.diagnostic_name], dst=out[self.tendency_name],
                           origin=(0, 0, 0), domain=(g.nx, g.ny, g.nz))


class FromTendencyToDiagnostic(_DomainComponent):
    """The tendency ``tendency_name`` copied (on the grid box) into the diagnostic ``diagnostic_name``."""

    kind = "t2d"

    def __init__(self, grid, tendency_name, diagnostic_name=None, **kwargs):
        super().__init__(grid, **kwargs)
        self.tendency_name = tendency_name
        self.diagnostic_name = diagnostic_name or "tendency_of_" + tendency_name
        self.diagnostic_names = (self.diagnostic_name,)
        self._stencil_copy = self.compile_stencil("copy")

    def array_call(self, tendencies, out):
        g = self.grid
        self._stencil_copy(src=tendencies[self.tendency_name], dst=out[self.diagnostic_name],
                           origin=(0, 0, 0), domain=(g.nx, g.ny, g.nz))


class AirPotentialTemperatureToDiagnostic(FromTendencyToDiagnostic):
    def __init__(self, grid, **kwargs):
        super().__init__(grid, "air_potential_temperature", "tendency_of_air_potential_temperature",
                         **kwargs)


class AirPotentialTemperatureToTendency(FromDiagnosticToTendency):
    def __init__(self, grid, **kwargs):
        super().__init__(grid, "tendency_of_air_potential_temperature", "air_potential_temperature",
                         **kwargs)


# ------------------------------------------------------------------------------ couplers
class ConcurrentCoupling(StencilFactory):
    """Tendencies of several components summed, diagnostics collected; ``serial`` policy: a
    component sees the diagnostics of the ones before it."""

    kind = "implicit"

    def __init__(self, *components, execution_policy="serial", backend="b200", backend_options=None,
                 storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        if execution_policy not in ("serial", "as_parallel"):
            raise ValueError(f"unknown execution policy {execution_policy!r}")
        self.components, self.execution_policy = components, execution_policy
        # concurrent_coupling_utils.py:L72-L83: the first component providing a tendency
        # overwrites the buffer, the later ones accumulate
        seen, self.overwrite_tendencies = [], []
        for c in components:
            self.overwrite_tendencies.append({n: n not in seen for n in c.tendency_names})
            seen += [n for n in c.tendency_names if n not in seen]
        self.tendency_names = tuple(seen)
        self.diagnostic_names = tuple(dict.fromkeys(n for c in components for n in c.diagnostic_names))
        self._shapes = {n: c.diagnostic_shape(n) for c in components for n in c.diagnostic_names}

    def diagnostic_shape(self, name):
        return self._shapes[name]

    def _allocate(self, state, out_tendencies, out_diagnostics):
        shape = None
        for n in self.tendency_names:
            if n not in out_tendencies:
                shape = shape or _state_shape(state)
                out_tendencies[n] = self.zeros(shape=state[n].shape if n in state else shape)
        for n in self.diagnostic_names:
            if n not in out_diagnostics:
                out_diagnostics[n] = self.zeros(shape=self._shapes[n])

    def __call__(self, state, timestep=None, *, out_tendencies=None, out_diagnostics=None,
                 overwrite_tendencies=None):
        out_tendencies = out_tendencies if out_tendencies is not None else {}
        out_diagnostics = out_diagnostics if out_diagnostics is not None else {}
        overwrite_tendencies = overwrite_tendencies or {}
        self._allocate(state, out_tendencies, out_diagnostics)
        serial = self.execution_policy == "serial"
        aux_state = dict(state) if serial else state
        for c, self_ow in zip(self.components, self.overwrite_tendencies):
            if c.kind == "diagnostic":
                c.array_call(aux_state, out_diagnostics)
            elif c.kind in ("tendency", "implicit"):
                ow = {n: self_ow[n] and overwrite_tendencies.get(n, True) for n in self_ow}
                if c.kind == "tendency":
                    c.array_call(aux_state, out_tendencies, out_diagnostics, ow)
                else:
                    c.array_call(aux_state, timestep, out_tendencies, out_diagnostics, ow)
            elif c.kind == "t2d":
                if serial:
                    c.array_call(out_tendencies, out_diagnostics)
            elif c.kind == "d2t":
                c.array_call(aux_state, out_tendencies)
            else:
                raise TypeError(f"cannot couple a component of kind {c.kind!r}")
            if serial:
                aux_state.update({n: out_diagnostics[n] for n in c.diagnostic_names})
        if "time" in state:
            out_tendencies["time"] = out_diagnostics["time"] = state["time"]
        return out_tendencies, out_diagnostics

    # the couplers call components through array_call
    def array_call(self, state, timestep, out_tendencies, out_diagnostics, overwrite_tendencies):
        self(state, timestep, out_tendencies=out_tendencies, out_diagnostics=out_diagnostics,
             overwrite_tendencies=overwrite_tendencies)


class TendencyStepper(StencilFactory):
    """``stepper(state, timestep) -> (diagnostics of the first stage, stepped fields)``."""

    # fraction of the timestep each stage advances the *initial* state by
    SCHEMES = {
        "forward_euler": (1.0,),            # forward_euler.py:L57-L72
        "rk2": (0.5, 1.0),                  # rk2.py:L60-L118
        "rk3ws": (1.0 / 3.0, 0.5, 1.0),     # rk3ws.py:L60-L157
    }

    @staticmethod
    def factory(scheme, *components, **kwargs):
        if scheme not in TendencyStepper.SCHEMES:
            raise ValueError(f"unknown (or out-of-scope) tendency stepper {scheme!r}")
        return TendencyStepper(scheme, *components, **kwargs)

    def __init__(self, scheme, *components, execution_policy="serial",
                 enforce_horizontal_boundary=False, backend="b200", backend_options=None,
                 storage_options=None):
        super().__init__(backend, backend_options or BackendOptions(), storage_options or StorageOptions())
        self.scheme, self._factors = scheme, self.SCHEMES[scheme]
        # steppers.py:L96-L107: several components -> one ConcurrentCoupling
        if len(components) > 1 or not isinstance(components[0], ConcurrentCoupling):
            self.prognostic = ConcurrentCoupling(
                *components, execution_policy=execution_policy, backend=backend,
                storage_options=self.storage_options)
        else:
            self.prognostic = components[0]
        self._hb = None
        if enforce_horizontal_boundary:  # steppers.py:L110-L126
            for c in components:
                self._hb = getattr(c, "horizontal_boundary", None)
                if self._hb is not None:
                    break
        self._increment, self._diagnostics = {}, {}
        # b200: a lone tendency component that can apply the stage update itself
        # (``array_call_stepped``: out = base + factor * tendency in the tendency kernel) spares
        # the round trip of the tendencies through memory.  Stages read the previous stage's
        # output, so a second set of stage buffers alternates with out_state (the last stage
        # writes out_state).  TB200_FUSED_STEP=0 keeps the generic tendencies + fma path.
        lone = components[0] if len(components) == 1 else None
        self._fused = (lone if lone is not None and getattr(lone, "kind", None) == "tendency"
                       and hasattr(lone, "array_call_stepped") and not lone.diagnostic_names
                       and os.environ.get("TB200_FUSED_STEP", "1") != "0" else None)
        self._stage_buffers = {}

    def _call_fused(self, state, timestep, out_state, names):
        dt = timestep.total_seconds()
        if len(self._factors) > 1:
            for n in names:
                if n not in self._stage_buffers:
                    self._stage_buffers[n] = self.zeros(shape=state[n].shape)
        nst = len(self._factors)
        cur = state
        for stage, c in enumerate(self._factors):
            dst = out_state if (nst - 1 - stage) % 2 == 0 else self._stage_buffers
            self._fused.array_call_stepped(cur, state, c * dt, dst)
            if self._hb is not None:
                self._hb.enforce_raw(dst, {n: {} for n in names})
            if stage < nst - 1:
                cur = dict(state)
                cur.update({n: dst[n] for n in names})

    def output_names(self, state):
        return tuple(n for n in self.prognostic.tendency_names if n in state)

    def __call__(self, state, timestep, *, out_diagnostics=None, out_state=None):
        out_diagnostics = out_diagnostics if out_diagnostics is not None else {}
        out_state = out_state if out_state is not None else {}
        names = self.output_names(state)
        for n in names:
            if n not in out_state:
                out_state[n] = self.zeros(shape=state[n].shape)
        if self._fused is not None and set(names) == set(self._fused.tendency_names):
            self._call_fused(state, timestep, out_state, names)
            if "time" in state:
                out_state["time"] = state["time"] + timestep
            return out_diagnostics, out_state
        dt = timestep.total_seconds()
        cur = state
        for stage, c in enumerate(self._factors):
            # the diagnostics returned are those of the first stage (rk2.py:L62-L71)
            diags = out_diagnostics if stage == 0 else self._diagnostics
            self.prognostic(cur, timestep, out_tendencies=self._increment, out_diagnostics=diags)
            # one launch per storage shape (a stepped field may be a one-level 2-D field), like
            # plugin._batch_dict_operator_fma does for the reference's DataArrayDictOperator.fma
            groups = {}
            for n in names:
                groups.setdefault(tuple(out_state[n].shape), []).append(n)
            for shape, ns in groups.items():
                stencils.fma_fields([out_state[n] for n in ns], [state[n] for n in ns],
                                    [self._increment[n] for n in ns], c * dt,
                                    origin=(0, 0, 0), domain=shape)
            if self._hb is not None:
                self._hb.enforce_raw(out_state, {n: {} for n in names})
            if stage < len(self._factors) - 1:
                # every other variable comes from the state (rk2.py:L88-L91)
                cur = dict(state)
                cur.update({n: out_state[n] for n in names})
                if "time" in state:
                    cur["time"] = state["time"] + c * timestep
        if "time" in state:
            out_state["time"] = state["time"] + timestep
        return out_diagnostics, out_state


class TimeIntegrationOptions:
    def __init__(self, component, scheme=None, substeps=1, enforce_horizontal_boundary=False,
                 **kwargs):
        self.component, self.scheme, self.substeps = component, scheme, substeps
        self.enforce_horizontal_boundary = enforce_horizontal_boundary
        self.kwargs = kwargs


class SequentialUpdateSplitting:
    """Components applied one after the other, each on the state the previous one left."""

    def __init__(self, *args):
        self._component_list = []
        for options in args:
            c = options.component
            if c.kind == "diagnostic":
                self._component_list.append(c)
            else:  # sequential_update_splitting.py:L113-L128
                if max(options.substeps, 1) > 1:
                    raise NotImplementedError("substepping")
                self._component_list.append(TendencyStepper.factory(
                    options.scheme or "forward_euler", c, execution_policy="serial",
                    enforce_horizontal_boundary=options.enforce_horizontal_boundary))
        self._out_diagnostics = [{} for _ in self._component_list]
        self._out_state = [{} for _ in self._component_list]

    def __call__(self, state, timestep):
        current_time = state.get("time")
        for idx, c in enumerate(self._component_list):
            diags, outs = self._out_diagnostics[idx], self._out_state[idx]
            if isinstance(c, TendencyStepper):
                c(state, timestep, out_diagnostics=diags, out_state=outs)
                update_swap(state, diags)
                update_swap(state, outs)
            else:
                for n in c.diagnostic_names:
                    if n not in diags:
                        diags[n] = c.zeros(shape=c.diagnostic_shape(n))
                c.array_call(state, diags)
                update_swap(state, diags)
            if current_time is not None:
                state["time"] = current_time
        if current_time is not None:
            state["time"] = current_time + timestep
