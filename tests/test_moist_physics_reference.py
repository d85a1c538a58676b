# -*- coding: utf-8 -*-
"""The moist physics pass of BASELINE configs[2] executed by the REFERENCE ITSELF (run in place
from /root/reference, numpy backend; skipped where the tree is absent) against the oracle's
``MoistIsentropicModel.physics``, bit for bit.

Everything that computes or orders is the reference's own code:
  * the eleven component objects are the reference's classes, built by their own constructors
    with the namelist's parameters (namelist_sus.py), and are entered through ``array_call``;
  * tendencies are summed / promoted by ``ConcurrentCoupling._call_serial`` with the overwrite
    flags of ``StaticOperator.get_overwrite_tendencies``;
  * stages are taken by ``ForwardEuler / RK2 / RK3WS._call`` with ``DataArrayDictOperator.fma``;
  * the chain is walked by ``SequentialUpdateSplitting.__call__`` with ``update_swap``.
What the private sympl fork would contribute -- the ``__call__`` wrappers that turn DataArray
dicts into raw-array dicts and allocate outputs -- is played by the small adapters below.  The
component list and its order are those of driver_namelist_sus.py:L184-L479 (a script; by reading).
"""
import importlib
import types
from datetime import timedelta

import numpy as np
import pytest

from oracle import moist_model as mm
from tests.golden import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference tree not mounted")

NX, NY, NZ, NB = 21, 19, 10, 3
SHAPE = (NX + 1, NY + 1, NZ + 1)


def ref(mod):
    refload.install_framework()
    return importlib.import_module(mod)


def units_of(name):
    from tasmania_b200.iox import UNITS

    base = name.replace("tendency_of_", "") if name not in UNITS else name
    return UNITS.get(name, UNITS.get(base, "1"))


def da(arr, name):
    return refload.DataArray(arr, None, ("x", "y", "z"), None, {"units": units_of(name)})


def raw(d):
    """DataArray dict -> raw-array dict (the time stamp rides along, as in sympl)."""
    return {n: (v if n == "time" else v.data) for n, v in d.items()}


class Properties:  # sympl's StaticComponentOperator, as far as the couplers use it
    def __init__(self, name):
        self.name = name

    @classmethod
    def factory(cls, name):
        return cls(name)

    def get_properties(self, component):
        return getattr(component, self.name, {})


def make_adapters(cc):
    """The sympl-side call wrappers around ``array_call``."""

    def fill(out, names, shape_of):
        for n in names:
            if n not in out:
                out[n] = da(np.zeros(shape_of(n)), n)

    class Diagnostic(cc.DiagnosticComponent):
        def __init__(self, comp, shape_of=lambda n: SHAPE):
            self.comp, self.shape_of = comp, shape_of
            self.diagnostic_properties = comp.diagnostic_properties
            self.tendency_properties = {}

        def __call__(self, state, out=None):
            out = out if out is not None else {}
            fill(out, self.diagnostic_properties, self.shape_of)
            self.comp.array_call(raw(state), raw(out))
            return out

    class Tendency(cc.TendencyComponent):
        implicit = False

        def __init__(self, comp, shape_of=lambda n: SHAPE):
            self.comp, self.shape_of = comp, shape_of
            self.tendency_properties = comp.tendency_properties
            self.diagnostic_properties = comp.diagnostic_properties

        def __call__(self, state, *args, out_tendencies=None, out_diagnostics=None,
                     overwrite_tendencies=None):
            if self.implicit and not args:
                raise TypeError("an ImplicitTendencyComponent takes the timestep")
            fill(out_tendencies, self.tendency_properties, self.shape_of)
            fill(out_diagnostics, self.diagnostic_properties, self.shape_of)
            self.comp.array_call(raw(state), *args, raw(out_tendencies), raw(out_diagnostics),
                                 overwrite_tendencies)

    class Implicit(Tendency, cc.ImplicitTendencyComponent):
        implicit = True

    return Diagnostic, Tendency, Implicit


def promoter(cls, grid, **props):
    copy_numpy = ref("tasmania.framework.subclasses.stencil_definitions.copy").copy_numpy

    class Promoter(cls):
        def __init__(self):
            pass

        def __call__(self, arrays, *, out=None):
            for n in list(self.tendency_properties) + list(self.diagnostic_properties):
                if n not in out:
                    out[n] = da(np.zeros(SHAPE), n)
            self.array_call(raw(arrays), raw(out))
            return out

    p = Promoter()
    p.__dict__.update(props)
    p._stencil_copy = lambda src, dst, origin, domain, **kw: copy_numpy(src, dst, origin=origin, domain=domain)
    p._backend_options = types.SimpleNamespace(exec_info=None, validate_args=False)
    Promoter.grid = grid
    Promoter.backend_options = property(lambda self: self._backend_options)
    return p


def reference_physics(model, domain):
    """The reference's physics suite of the moist benchmark as a callable ``physics(state, dt)`` on
    a DataArray state dict (in place)."""
    from tests.golden import generate_golden as gg

    pt = model.pt
    cc = ref("tasmania.framework.concurrent_coupling")
    ccu = ref("tasmania.framework.concurrent_coupling_utils")
    sus = ref("tasmania.framework.sequential_update_splitting")
    xr = ref("tasmania.utils.xarrayx")
    cc.StaticComponentOperator = Properties
    ccu.StaticOperator.tendency_operator = Properties("tendency_properties")
    steppers = {"forward_euler": ref("tasmania.framework.subclasses.tendency_steppers.forward_euler").ForwardEuler,
                "rk2": ref("tasmania.framework.subclasses.tendency_steppers.rk2").RK2,
                "rk3ws": ref("tasmania.framework.subclasses.tendency_steppers.rk3ws").RK3WS}
    grid = domain.numerical_grid
    opts = ref("tasmania.framework.options")
    kw = lambda: dict(enable_checks=False, backend="numpy", backend_options=opts.BackendOptions(),  # noqa: E731
                      storage_shape=SHAPE, storage_options=opts.StorageOptions())
    idg = ref("tasmania.isentropic.physics.diagnostics")
    co = ref("tasmania.isentropic.physics.coriolis")
    hsm = ref("tasmania.isentropic.physics.horizontal_smoothing")
    ref("tasmania.dwarfs.subclasses.horizontal_smoothers.second_order")
    tu = ref("tasmania.isentropic.physics.turbulence")
    iu = ref("tasmania.isentropic.utils")
    ke = ref("tasmania.physics.microphysics.kessler")
    ut = ref("tasmania.physics.microphysics.utils")
    ref("tasmania.physics.microphysics.sedimentation_fluxes.second_order")
    va = ref("tasmania.isentropic.physics.vertical_advection")
    ref("tasmania.isentropic.dynamics.subclasses.minimal_vertical_fluxes.third_order_upwind")
    g = gg.da
    Diagnostic, Tendency, Implicit = make_adapters(cc)
    sec2d = lambda n: (SHAPE[0], SHAPE[1], 1)  # noqa: E731
    prop = {"dims": ("x", "y", "z"), "units": "K s^-1"}
    t2d = promoter(iu.AirPotentialTemperatureToDiagnostic, grid, diagnostic_properties={mm.W: prop},
                   tendency_properties={})
    d2t = promoter(iu.AirPotentialTemperatureToTendency, grid, tendency_properties={mm.THETA: prop},
                   diagnostic_properties={})
    rfv = Diagnostic(ke.KesslerFallVelocity(domain, "numerical", **kw()))
    chain = [  # driver_namelist_sus.py:L184-L479 with namelist_sus.py:L33-L141
        Diagnostic(idg.IsentropicDiagnostics(domain, "numerical", True, g(pt, "Pa"), **kw())),
        ("rk2", [Tendency(co.IsentropicConservativeCoriolis(domain, grid_type="numerical",
                                                             coriolis_parameter=None, **kw()))]),
        Diagnostic(hsm.IsentropicHorizontalSmoothing(
            domain, "second_order", 1.0, 1.0, 0, moist=True, smooth_moist_coeff=1.0,
            smooth_moist_coeff_max=1.0, smooth_moist_damp_depth=0, **kw())),
        ("rk2", [Tendency(tu.IsentropicSmagorinsky(domain, 0.18, **kw()))]),
        Diagnostic(idg.IsentropicVelocityComponents(domain, **kw())),
        ("rk2", [Tendency(ke.KesslerMicrophysics(
            domain, "numerical", air_pressure_on_interface_levels=True,
            tendency_of_air_potential_temperature_in_diagnostics=False, rain_evaporation=True,
            autoconversion_threshold=g(0.1, "g kg^-1"), autoconversion_rate=g(0.001, "s^-1"),
            collection_rate=g(2.2, "s^-1"), **kw())), t2d]),
        ("rk2", [d2t, Tendency(ke.KesslerSaturationAdjustmentPrognostic(
            domain, grid_type="numerical", air_pressure_on_interface_levels=True,
            saturation_rate=g(0.025, "s^-1"), **kw())), t2d]),
        ("rk3ws", [Tendency(va.IsentropicVerticalAdvection(
            domain, flux_scheme="third_order_upwind", moist=True,
            tendency_of_air_potential_temperature_on_interface_levels=False, **kw()))]),
        ("rk3ws", [rfv, Implicit(ke.KesslerSedimentation(
            domain, "numerical", sedimentation_flux_scheme="second_order_upwind", **kw()))]),
        ("forward_euler", [rfv, Implicit(ut.Precipitation(domain, "numerical", **kw()), sec2d)]),
    ]
    op = xr.DataArrayDictOperator(backend="numpy")

    class Stepper:
        """sympl's TendencyStepper.__call__: allocate, then the reference's own ``_call``."""

        _enforce_hb = False

        def __init__(self, scheme, components):
            self.cls, self._dict_op, self._increment, self._diagnostics = steppers[scheme], op, None, None
            self.coupler = types.SimpleNamespace(
                components=tuple(components), execution_policy="serial",
                allowed_diagnostic_type=cc.ConcurrentCoupling.allowed_diagnostic_type,
                allowed_tendency_type=cc.ConcurrentCoupling.allowed_tendency_type)
            self.coupler.overwrite_tendencies = ccu.StaticOperator.get_overwrite_tendencies(self.coupler)
            self._stepper_operator = self

        def get_increment(self, state, timestep, out_increment=None, out_diagnostics=None):
            tnd = out_increment if out_increment is not None else {}
            diag = out_diagnostics if out_diagnostics is not None else {}
            cc.ConcurrentCoupling._call_serial(self.coupler, state, timestep, tnd, diag, {})
            tnd["time"] = diag["time"] = state["time"]
            return tnd, diag

        def __call__(self, state, timestep, out_diagnostics=None, out_state=None):
            names = [n for c in self.coupler.components for n in c.tendency_properties if n in state]
            self.output_properties = {n: {"units": state[n].attrs["units"]} for n in names}
            out_state = out_state if out_state is not None else {}
            for n in names:
                if n not in out_state:
                    out_state[n] = da(np.zeros(SHAPE), n)
            out_diagnostics = out_diagnostics if out_diagnostics is not None else {}
            return self.cls._call(self, state, timestep, out_diagnostics, out_state)

    components = [c if not isinstance(c, tuple) else Stepper(*c) for c in chain]
    me = types.SimpleNamespace(
        _component_list=components, _substeps=[1] * len(components),
        _out_diagnostics=[None] * len(components), _out_state=[None] * len(components), _dict_op=op,
        allowed_diagnostic_type=sus.SequentialUpdateSplitting.allowed_diagnostic_type)
    return lambda state, dt: sus.SequentialUpdateSplitting.__call__(me, state, dt)


def reference_dycore(model, domain):
    """The reference's moist dynamical core as a callable ``dycore(raw_state, dt) -> raw outputs``:
    ``IsentropicDynamicalCore.stage_array_call_moist`` (dycore.py:L723-L843, unbound) over the
    reference's RK3WSSI prognostic, Rayleigh damper (last stage only, namelist_sus.py:L102),
    HorizontalVelocity and WaterConstituent, stages chained as framework/dycore.py:L455-L458."""
    from tests.golden import generate_golden as gg

    dyc = ref("tasmania.isentropic.dynamics.dycore")
    ref("tasmania.isentropic.dynamics.subclasses.prognostics.utils")
    ref("tasmania.isentropic.dynamics.subclasses.prognostics.rk3ws_si")
    ref("tasmania.isentropic.dynamics.subclasses.minimal_horizontal_fluxes.fifth_order_upwind")
    ref("tasmania.dwarfs.subclasses.vertical_dampers.rayleigh")
    prog = ref("tasmania.isentropic.dynamics.prognostic")
    vd = ref("tasmania.dwarfs.vertical_damping")
    dd = ref("tasmania.dwarfs.diagnostics")
    opts = ref("tasmania.framework.options")
    bo, so = opts.BackendOptions, opts.StorageOptions
    grid, hb = domain.numerical_grid, domain.horizontal_boundary
    prognostic = prog.IsentropicPrognostic.factory(
        "rk3ws_si", "fifth_order_upwind", domain, True, backend="numpy", backend_options=bo(),
        storage_shape=SHAPE, storage_options=so(), pt=gg.da(model.pt, "Pa"), eps=0.5)
    outnames = (mm.S, mm.SU, mm.U, mm.SV, mm.V, mm.QV, mm.QC, mm.QR)
    me = types.SimpleNamespace(
        _water_constituent=dd.WaterConstituent(grid, clipping=True, backend="numpy", backend_options=bo(),
                                               storage_options=so()),
        **{f"_{q}_{t}": np.zeros(SHAPE) for q in ("sqv", "sqc", "sqr") for t in ("now", "int", "new")},
        horizontal_boundary=hb, output_properties={k: {"units": units_of(k)} for k in outnames},
        _damp=True, _damp_at_every_stage=False, stages=prognostic.stages, _prognostic=prognostic,
        _damper=vd.VerticalDamping.factory("rayleigh", grid, 4, 5e-4, backend="numpy", backend_options=bo(),
                                           storage_shape=SHAPE, storage_options=so()),
        _velocity_components=dd.HorizontalVelocity(grid, staggering=True, backend="numpy",
                                                   backend_options=bo(), storage_options=so()),
        _s_ref=np.zeros(SHAPE), _su_ref=np.zeros(SHAPE), _sv_ref=np.zeros(SHAPE),
        _s_now=None, _su_now=None, _sv_now=None)
    outs = [{k: np.zeros(SHAPE) for k in outnames} for _ in range(prognostic.stages)]

    def call(cur, dt):
        st_in = cur
        for stage in range(prognostic.stages):
            dyc.IsentropicDynamicalCore.stage_array_call_moist(me, stage, st_in, {}, dt, outs[stage])
            st_in = dict(outs[stage])
            st_in.setdefault(mm.MTG, cur[mm.MTG])
        return {k: outs[-1][k].copy() for k in outnames}

    return call


def _setup(nsteps_before):
    from tests.golden import generate_golden as gg
    from tests.test_moist_model_oracle import build

    model, ost = build(NX, NY, NZ, max_height=500.0)  # the terrain gg._make_domain builds
    dt = timedelta(seconds=5)
    for _ in range(nsteps_before):
        ost = model.step(ost, dt)
    domain = gg._make_domain(NX, NY, NZ, "relaxed", NB, {"nr": 6}, topo_time=60.0)
    domain.numerical_grid.update_topography(model.nstep * dt)
    domain.horizontal_boundary.reference_state = {
        n: da(v.copy(), n) for n, v in model.hb.reference_state.items()}
    return model, ost, domain, dt


def test_reference_physics_pass_equals_oracle():
    """One pass of the physics suite on an evolved state (cloud, rain, latent heating non-zero)."""
    model, ost, domain, dt = _setup(3)
    physics = reference_physics(model, domain)
    rstate = {n: da(v.copy(), n) for n, v in ost.items() if n != "time"}
    rstate["time"] = ost["time"]
    want = dict(ost)
    with np.errstate(divide="ignore", invalid="ignore"):
        physics(rstate, dt)
        model.physics(want, dt)
    assert rstate["time"] == want["time"] and set(rstate) == set(want)
    for n, v in want.items():
        if n != "time":
            np.testing.assert_array_equal(rstate[n].data, v, err_msg=n)
    box = (slice(0, NX), slice(0, NY), slice(0, NZ))
    assert float(want[mm.QR][box].max()) > 1e-5 and float(np.abs(want[mm.W][box]).max()) > 0.0
    assert float(want[mm.ACCPREC].max()) > 0.0


def test_reference_model_steps_equal_oracle():
    """Five full steps of the benchmark loop (driver_namelist_sus.py:L490-L512): the reference's
    moist dynamical core and physics suite, run in place, against ``MoistIsentropicModel.step``."""
    model, ost, domain, dt = _setup(0)
    physics, dycore = reference_physics(model, domain), reference_dycore(model, domain)
    grid = domain.numerical_grid
    rstate = {n: da(v.copy(), n) for n, v in ost.items() if n != "time"}
    rstate["time"] = ost["time"]
    with np.errstate(divide="ignore", invalid="ignore"):
        for step in range(5):
            grid.update_topography((step + 1) * dt)
            out = dycore(raw(rstate), dt)
            for n, v in out.items():  # the fields the dycore does not return are carried over (L503)
                rstate[n] = da(v, n)
            physics(rstate, dt)
            ost = model.step(ost, dt)
            assert rstate["time"] == ost["time"]
            for n, v in ost.items():
                if n != "time":
                    np.testing.assert_array_equal(rstate[n].data, v, err_msg=f"step {step}: {n}")
    assert float(np.abs(ost[mm.SV]).max()) > 1.0
