# -*- coding: utf-8 -*-
"""Run by tests/test_plugin_reference.py in a subprocess.  The UNMODIFIED reference physics components of
the moist benchmark (driver_namelist_sus.py:L184-L473), constructed with backend="b200" through the
plugin, have their ``array_call`` run against the recording C-ABI stub (tests/abi_stub.py; storages on
the host); the ABI calls each one issues -- kernel, scalars, boxes, canonical buffer ids with shapes
and strides (sliced views included) -- must be the ones the b200 host mirror of the same component
issues on the same state.  This is north_star's "the sympl TendencyComponent / DiagnosticComponent
classes ... work unchanged" up to the library boundary, and it pins the mirrors' wiring.
Masked: buffers whose rank is a representational choice of the mirror (coefficients stored at
their true rank and broadcast through zero strides).
"""
import os
import sys
from datetime import datetime, timedelta

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import refload  # noqa: E402

refload.install_framework()

import generate_golden as gg  # noqa: E402
import tasmania_b200 as tb  # noqa: E402
from tasmania_b200 import coupling, isentropic_physics, microphysics, plugin  # noqa: E402
from tests import helpers as hp  # noqa: E402
from tests.abi_stub import stubbed_library  # noqa: E402

plugin.install()

NX, NY, NZ, NB = 17, 15, 8, 3
SHAPE = (NX + 1, NY + 1, NZ + 1)
S, SU, SV = gg.S, gg.SU, gg.SV
QV, QC, QR = gg.MFWV, gg.MFCW, gg.MFPW
THETA, W, VT = "air_potential_temperature", "tendency_of_air_potential_temperature", "raindrop_fall_velocity"
DT = timedelta(seconds=5)


def is_field(x):
    return isinstance(x, tuple) and len(x) == 3 and isinstance(x[1], tuple) and isinstance(x[0], int)


def reduce(trace, mask, keep=None):
    ids, out = {}, []
    for name, desc in trace:
        if keep is not None and name not in keep:
            continue
        row = [name]
        for pos, x in enumerate(desc):
            if pos in mask.get(name, ()):
                row.append("masked")
            elif is_field(x):
                row.append(("field", ids.setdefault(x[0], len(ids)), x[1], x[2]))
            else:
                row.append(x)
        out.append(tuple(row))
    return out


def record(stub, fn):
    stub.trace = []
    fn()
    trace, stub.trace = stub.trace, None
    return trace


def upload(np_state):
    st = {n: tb.as_storage(v) for n, v in np_state.items()}
    st["time"] = datetime(1992, 2, 20)
    return st


def buffers(names, shape=SHAPE):
    return {n: tb.zeros(shape) for n in names}


with stubbed_library() as stub:
    grid, np_state = hp.moist_case(NX, NY, NZ)
    rng = np.random.default_rng(8)
    np_state[W] = rng.standard_normal(SHAPE) * 1e-3
    np_state[VT] = rng.uniform(0.0, 5.0, SHAPE)
    domain = gg._make_domain(NX, NY, NZ, "relaxed", NB, {"nr": 6}, topo_time=60.0)
    opts = refload.load("tasmania.framework.options")
    kw = dict(enable_checks=False, backend="b200", backend_options=None, storage_shape=SHAPE,
              storage_options=opts.StorageOptions())

    def ref_kw():
        return dict(kw, backend_options=opts.BackendOptions())

    ke = refload.load("tasmania.physics.microphysics.kessler")
    ut = refload.load("tasmania.physics.microphysics.utils")
    for m in ("first_order", "second_order"):
        refload.load("tasmania.physics.microphysics.sedimentation_fluxes." + m)
    va = refload.load("tasmania.isentropic.physics.vertical_advection")
    for m in ("upwind", "centered", "third_order_upwind", "fifth_order_upwind"):
        refload.load("tasmania.isentropic.dynamics.subclasses.minimal_vertical_fluxes." + m)
    co = refload.load("tasmania.isentropic.physics.coriolis")
    tu = refload.load("tasmania.isentropic.physics.turbulence")
    da = gg.da
    all_true = lambda names: {n: True for n in names}  # noqa: E731
    mkw = dict(storage_shape=SHAPE)
    sec2d = (SHAPE[0], SHAPE[1], 1)
    checked = []

    def compare(label, ref_call, mirror_call, mask=None, keep=None):
        a = reduce(record(stub, ref_call), mask or {}, keep)
        b = reduce(record(stub, mirror_call), mask or {}, keep)
        assert a and len(a) == len(b), (label, len(a), len(b))
        for n, (p, q) in enumerate(zip(a, b)):
            assert p == q, (label, n, p, q)
        checked.append(label)

    # ---- Kessler microphysics
    names = (QC, QR, QV, THETA)
    r = ke.KesslerMicrophysics(
        domain, "numerical", air_pressure_on_interface_levels=True,
        tendency_of_air_potential_temperature_in_diagnostics=False, rain_evaporation=True,
        autoconversion_threshold=da(0.1, "g kg^-1"), autoconversion_rate=da(0.001, "s^-1"),
        collection_rate=da(2.2, "s^-1"), **ref_kw())
    m = microphysics.KesslerMicrophysics(grid, autoconversion_threshold=0.1e-3, autoconversion_rate=0.001,
                                         collection_rate=2.2, **mkw)
    assert set(m.tendency_names) == set(r.tendency_properties) == set(names)
    sr, sm, tr_, tm = upload(np_state), upload(np_state), buffers(names), buffers(names)
    compare("KesslerMicrophysics", lambda: r.array_call(sr, tr_, {}, all_true(names)),
            lambda: m.array_call(sm, tm, {}, all_true(names)))

    # ---- saturation adjustment (prognostic)
    names = (QV, QC, THETA)
    r = ke.KesslerSaturationAdjustmentPrognostic(
        domain, grid_type="numerical", air_pressure_on_interface_levels=True,
        saturation_rate=da(0.025, "s^-1"), **ref_kw())
    m = microphysics.KesslerSaturationAdjustmentPrognostic(grid, saturation_rate=0.025, **mkw)
    assert set(m.tendency_names) == set(r.tendency_properties)
    ow = {QV: True, QC: True, THETA: False}
    tr_, tm = buffers(names), buffers(names)
    compare("KesslerSaturationAdjustmentPrognostic", lambda: r.array_call(sr, tr_, {}, ow),
            lambda: m.array_call(sm, tm, {}, ow))

    # ---- fall velocity, sedimentation, precipitation
    r = ke.KesslerFallVelocity(domain, "numerical", **ref_kw())
    m = microphysics.KesslerFallVelocity(grid, **mkw)
    assert set(m.diagnostic_names) == set(r.diagnostic_properties)
    dr, dm = buffers((VT,)), buffers((VT,))
    compare("KesslerFallVelocity", lambda: r.array_call(sr, dr), lambda: m.array_call(sm, dm))
    r = ke.KesslerSedimentation(domain, "numerical", sedimentation_flux_scheme="second_order_upwind",
                                **ref_kw())
    m = microphysics.KesslerSedimentation(grid, sedimentation_flux_scheme="second_order_upwind", **mkw)
    assert set(m.tendency_names) == set(r.tendency_properties)
    tr_, tm = buffers((QR,)), buffers((QR,))
    compare("KesslerSedimentation", lambda: r.array_call(sr, DT, tr_, {}, {QR: True}),
            lambda: m.array_call(sm, DT, tm, {}, {QR: True}))
    r = ut.Precipitation(domain, "numerical", **ref_kw())
    m = microphysics.Precipitation(grid, **mkw)
    assert set(m.diagnostic_names) == set(r.diagnostic_properties)
    dr = buffers(("precipitation", "accumulated_precipitation"), sec2d)
    dm = buffers(("precipitation", "accumulated_precipitation"), sec2d)
    compare("Precipitation", lambda: r.array_call(sr, DT, {}, dr, {}),
            lambda: m.array_call(sm, DT, {}, dm, {}))

    # ---- vertical advection (moist, third-order upwind, w on the main levels)
    names = (S, SU, SV, QV, QC, QR)
    r = va.IsentropicVerticalAdvection(
        domain, flux_scheme="third_order_upwind", moist=True,
        tendency_of_air_potential_temperature_on_interface_levels=False, **ref_kw())
    m = isentropic_physics.IsentropicVerticalAdvection(grid, flux_scheme="third_order_upwind", moist=True,
                                                       **mkw)
    assert set(m.tendency_names) == set(r.tendency_properties)
    tr_, tm = buffers(names), buffers(names)
    compare("IsentropicVerticalAdvection", lambda: r.array_call(sr, tr_, {}, all_true(names)),
            lambda: m.array_call(sm, tm, {}, all_true(names)))

    # ---- Coriolis, Smagorinsky
    names = (SU, SV)
    r = co.IsentropicConservativeCoriolis(domain, grid_type="numerical", coriolis_parameter=None, **ref_kw())
    m = isentropic_physics.IsentropicConservativeCoriolis(grid, NB)
    tr_, tm = buffers(names), buffers(names)
    compare("IsentropicConservativeCoriolis", lambda: r.array_call(sr, tr_, {}, all_true(names)),
            lambda: m.array_call(sm, tm, {}, all_true(names)))
    r = tu.IsentropicSmagorinsky(domain, 0.18, **ref_kw())
    m = isentropic_physics.IsentropicSmagorinsky(grid, NB, 0.18)
    tr_, tm = buffers(names), buffers(names)
    compare("IsentropicSmagorinsky", lambda: r.array_call(sr, tr_, {}, all_true(names)),
            lambda: m.array_call(sm, tm, {}, all_true(names)))

    # ---- the diagnostic components.  The reference smooths with one launch + four rim copies
    # (first_order.py:L77-L110) where the mirror's kernel copies the rim itself, sets the
    # outermost velocity faces by slice assignment where the mirror launches a kernel, and keeps
    # theta / topography / gamma as 3-D storages: only the shared kernels are compared, those
    # coefficient buffers masked.
    hs = refload.load("tasmania.isentropic.physics.horizontal_smoothing")
    for mname in ("first_order", "second_order", "third_order"):
        refload.load("tasmania.dwarfs.subclasses.horizontal_smoothers." + mname)
    names = (S, SU, SV, QV, QC, QR)
    r = hs.IsentropicHorizontalSmoothing(domain, "second_order", 1.0, 1.0, 0, moist=True,
                                         smooth_moist_coeff=1.0, smooth_moist_coeff_max=1.0,
                                         smooth_moist_damp_depth=0, **ref_kw())
    m = coupling.IsentropicHorizontalSmoothing(grid, NB, "second_order", 1.0, 1.0, 0, moist=True,
                                               smooth_moist_coeff=1.0, smooth_moist_coeff_max=1.0,
                                               smooth_moist_damp_depth=0, **mkw)
    assert set(m.diagnostic_names) == set(r.diagnostic_properties)
    dr, dm = buffers(names), buffers(names)
    compare("IsentropicHorizontalSmoothing", lambda: r.array_call(sr, dr), lambda: m.array_call(sm, dm),
            mask={"tb200_smoothing": {2, 4}}, keep=("tb200_smoothing",))  # gamma rank, rim_copy flag

    idg = refload.load("tasmania.isentropic.physics.diagnostics")
    pt = float(np_state["air_pressure_on_interface_levels"][0, 0, 0])
    names = ("air_pressure_on_interface_levels", "exner_function_on_interface_levels",
             "height_on_interface_levels", "montgomery_potential", "air_density", "air_temperature")
    r = idg.IsentropicDiagnostics(domain, "numerical", True, da(pt, "Pa"), **ref_kw())
    m = coupling.IsentropicDiagnostics(grid, True, pt, **mkw)
    assert set(m.diagnostic_names) == set(r.diagnostic_properties) == set(names)
    dr, dm = buffers(names), buffers(names)
    compare("IsentropicDiagnostics", lambda: r.array_call(sr, dr), lambda: m.array_call(sm, dm),
            mask={"tb200_diagnostic_variables": {0, 1}, "tb200_density_and_temperature": {0}},
            keep=("tb200_diagnostic_variables", "tb200_density_and_temperature"))

    hb = domain.horizontal_boundary
    from tasmania_b200.iox import UNITS  # noqa: E402

    hb.reference_state = {k: refload.DataArray(v, None, None, None, {"units": UNITS[k]})
                          for k, v in upload(np_state).items() if k != "time"}
    from tasmania_b200.boundary import Relaxed  # noqa: E402

    mhb = Relaxed(NX, NY, NZ, NB, nr=6)
    mhb.reference_state = upload(np_state)
    names = ("x_velocity_at_u_locations", "y_velocity_at_v_locations")
    r = idg.IsentropicVelocityComponents(domain, **ref_kw())
    m = coupling.IsentropicVelocityComponents(grid, mhb, **mkw)
    assert set(m.diagnostic_names) == set(r.diagnostic_properties)
    dr, dm = buffers(names), buffers(names)
    compare("IsentropicVelocityComponents", lambda: r.array_call(sr, dr), lambda: m.array_call(sm, dm),
            keep=("tb200_velocity",))

    # ---- horizontal diffusion as a tendency component (fourth order, dry + water species)
    hdf = refload.load("tasmania.isentropic.physics.horizontal_diffusion")
    for mname in ("second_order", "fourth_order"):
        refload.load("tasmania.dwarfs.subclasses.horizontal_diffusers." + mname)
    names = (S, SU, SV, QV, QC, QR)
    r = hdf.IsentropicHorizontalDiffusion(
        domain, "fourth_order", da(0.4, "s^-1"), da(0.9, "s^-1"), 4, moist=True,
        diffusion_moist_coeff=da(0.2, "s^-1"), diffusion_moist_coeff_max=da(0.5, "s^-1"),
        diffusion_moist_damp_depth=3, **ref_kw())
    m = coupling.IsentropicHorizontalDiffusion(grid, NB, "fourth_order", 0.4, 0.9, 4, moist=True,
                                               diffusion_moist_coeff=0.2, diffusion_moist_coeff_max=0.5,
                                               diffusion_moist_damp_depth=3, **mkw)
    assert set(m.tendency_names) == set(r.tendency_properties)
    np.testing.assert_array_equal(tb.to_numpy(m._core._gamma)[0, 0], tb.to_numpy(r._core._gamma)[0, 0])
    np.testing.assert_array_equal(tb.to_numpy(m._core_moist._gamma)[0, 0], tb.to_numpy(r._core_moist._gamma)[0, 0])
    ow = {n: (n != SU) for n in names}  # one accumulating field
    tr_, tm = buffers(names), buffers(names)
    compare("IsentropicHorizontalDiffusion", lambda: r.array_call(sr, tr_, {}, ow),
            lambda: m.array_call(sm, tm, {}, ow), mask={"tb200_diffusion": {2}})  # gamma rank

    # ---- the Burgers dwarf (BASELINE configs[0]): the reference's RK3WS stepper, third-order advection
    bst = refload.load("tasmania.burgers.dynamics.stepper")
    refload.load("tasmania.burgers.dynamics.subclasses.stepper.rk3ws")
    refload.load("tasmania.burgers.dynamics.subclasses.advection.third_order")
    from tasmania_b200.burgers import BurgersStepper  # noqa: E402
    from tasmania_b200.grid import Grid  # noqa: E402

    bnx, bny, bnb = 21, 19, 2
    bdomain = gg._make_domain(bnx, bny, 1, "relaxed", bnb, {"nr": 6}, topo=False)
    r = bst.BurgersStepper.factory("rk3ws", bdomain.numerical_grid.grid_xy, bnb, "third_order",
                                   backend="b200", backend_options=opts.BackendOptions(),
                                   storage_options=opts.StorageOptions())
    bgrid = Grid((-176.0, 176.0), bnx, (-176.0, 176.0), bny, (400.0, 280.0), 1, units_to_m=1e3)
    m = BurgersStepper.factory("rk3ws", bgrid, bnb, "third_order")
    assert r.stages == m.stages == 3
    vel = {n: rng.standard_normal((bnx, bny, 1)) for n in ("x_velocity", "y_velocity")}

    def burgers_step(stepper):
        state = {n: tb.as_storage(v) for n, v in vel.items()}
        state["time"] = datetime(2000, 1, 1)
        outs = [{n: tb.zeros((bnx, bny, 1)) for n in vel} for _ in range(3)]
        cur = state
        for stage in range(3):
            stepper(stage, cur, {}, DT, outs[stage])
            cur = outs[stage]
        assert cur["time"] == datetime(2000, 1, 1) + DT

    compare("BurgersStepper", lambda: burgers_step(r), lambda: burgers_step(m))

print("REF-COMPONENTS-OK", len(checked), " ".join(checked))
