# -*- coding: utf-8 -*-
"""Wiring of the moist physics components pinned on the REFERENCE run in place (skipped where
/root/reference is absent): the reference's own ``array_call`` methods of KesslerMicrophysics,
KesslerSaturationAdjustmentPrognostic, KesslerFallVelocity, KesslerSedimentation and
Precipitation (src/tasmania/physics/microphysics/{kessler,utils}.py), called unbound on a stand-in
``self`` that carries the reference's own numpy stencil with the externals the class injects,
against the tendency providers of the oracle's moist model (oracle/moist_model.py) on the same
evolved model state: which state field feeds which stencil argument, on which box, with which
overwrite flags -- bit for bit.  (The stencils themselves are pinned in tests/test_oracle_golden.py.)
"""
import types
from datetime import timedelta

import numpy as np
import pytest

from oracle import moist_model as mm
from tests.golden import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference tree not mounted")

RD, RV, CP, LHVW, RHOW = 287.05, 461.52, 1004.0, 2.5e6, 1.0e3


@pytest.fixture(scope="module")
def evolved():
    from tests.test_moist_model_oracle import build

    model, st = build(21, 19, 10)
    for _ in range(4):
        st = model.step(st, timedelta(seconds=5))
    return model, st


def stand_in(model, definition, externals, **attrs):
    stencil = refload.numpy_stencil(definition, externals)
    g = model.g
    return types.SimpleNamespace(
        grid=types.SimpleNamespace(nx=g.nx, ny=g.ny, nz=g.nz),
        backend_options=types.SimpleNamespace(exec_info=None, validate_args=False),
        _stencil=lambda exec_info=None, validate_args=None, **kw: stencil(**kw), **attrs)


def modules():
    refload.install_framework()
    return (refload.load("tasmania.physics.microphysics.kessler"),
            refload.load("tasmania.physics.microphysics.utils"),
            refload.load("tasmania.framework.subclasses.subroutine_definitions.generics"))


def test_kessler_microphysics_array_call(evolved):
    model, st = evolved
    ke, _, gen = modules()
    ext = {"air_pressure_on_interface_levels": True, "beta": RD / RV, "e": np.exp(1), "lhvw": LHVW,
           "rain_evaporation": True, "set_output": gen.set_output_numpy}
    me = stand_in(model, ke.KesslerMicrophysics._kessler_numpy, ext, _a=model.a, _k1=model.k1,
                  _k2=model.k2, _air_pressure_on_interface_levels=True, _rain_evaporation=True,
                  _pttd=False, _placeholder=model.z())
    names = (mm.QV, mm.QC, mm.QR, mm.THETA)
    tnd = {n: np.full(model.shape, 7.0) for n in names}  # stale content must be overwritten
    ke.KesslerMicrophysics.array_call(me, st, tnd, {}, {n: True for n in names})
    want, diag = model._kessler(st)
    box = (slice(0, model.g.nx), slice(0, model.g.ny), slice(0, model.g.nz))
    for n in names:
        np.testing.assert_array_equal(tnd[n][box], want[n][box], err_msg=n)
    assert float(np.abs(want[mm.QR][box]).max()) > 0.0 and float(np.abs(want[mm.THETA][box]).max()) > 0.0
    np.testing.assert_array_equal(diag[mm.W][box], want[mm.THETA][box])


def test_saturation_adjustment_array_call_accumulates_on_the_promoted_heating(evolved):
    model, st = evolved
    ke, _, gen = modules()
    ext = {"air_pressure_on_interface_levels": True, "beta": RD / RV, "cp": CP, "e": np.exp(1),
           "lhvw": LHVW, "rv": RV, "set_output": gen.set_output_numpy}
    me = stand_in(model, ke.KesslerSaturationAdjustmentPrognostic._saturation_prognostic_numpy, ext,
                  _apoil=True, _sr=model.sr)
    g = model.g
    tnd = {n: np.full(model.shape, 7.0) for n in (mm.QV, mm.QC)}
    tnd[mm.THETA] = model.z()
    tnd[mm.THETA][: g.nx, : g.ny, : g.nz] = st[mm.W][: g.nx, : g.ny, : g.nz]  # what d2t copied in
    ke.KesslerSaturationAdjustmentPrognostic.array_call(
        me, st, tnd, {}, {mm.QV: True, mm.QC: True, mm.THETA: False})
    want, _ = model._saturation(st)
    box = (slice(0, g.nx), slice(0, g.ny), slice(0, g.nz))
    for n in (mm.QV, mm.QC, mm.THETA):
        np.testing.assert_array_equal(tnd[n][box], want[n][box], err_msg=n)
    assert float(np.abs(want[mm.QC][box]).max()) > 0.0


def test_fall_velocity_sedimentation_and_precipitation_array_calls(evolved):
    model, st = evolved
    ke, ut, gen = modules()
    g = model.g
    box = (slice(0, g.nx), slice(0, g.ny), slice(0, g.nz))
    # ---- fall velocity (slab broadcast of the surface density, kessler.py:L1169)
    me = stand_in(model, ke.KesslerFallVelocity._fall_velocity_numpy, {}, _in_rho_s=model.z())
    out = {mm.VT: model.z()}
    ke.KesslerFallVelocity.array_call(me, st, out)
    np.testing.assert_array_equal(out[mm.VT][box], model._fall_velocity(st)[box])
    # ---- sedimentation, second-order upwind flux
    sf = refload.load("tasmania.physics.microphysics.sedimentation_fluxes.second_order").SecondOrderUpwind
    ext = {"set_output": gen.set_output_numpy, "sflux": sf.call_numpy, "sflux_extent": sf.nb}
    me = stand_in(model, ke.KesslerSedimentation._sedimentation_numpy, ext)
    st_vt = dict(st)
    st_vt[mm.VT] = out[mm.VT]
    tnd = {mm.QR: np.full(model.shape, 7.0)}
    ke.KesslerSedimentation.array_call(me, st_vt, timedelta(seconds=5), tnd, {}, {mm.QR: True})
    want, diag = model._sedimentation(st)
    np.testing.assert_array_equal(tnd[mm.QR][box], want[mm.QR][box])
    np.testing.assert_array_equal(diag[mm.VT][box], out[mm.VT][box])
    assert float(np.abs(want[mm.QR][box]).max()) > 0.0
    # ---- precipitation on the surface slabs (utils.py:L262-L281)
    me = stand_in(model, ut.Precipitation._accumulated_precipitation_numpy, {"rhow": RHOW})
    shape2d = (model.shape[0], model.shape[1], 1)
    diags = {mm.PREC: np.zeros(shape2d), mm.ACCPREC: np.zeros(shape2d)}
    ut.Precipitation.array_call(me, st_vt, timedelta(seconds=5), {}, diags, {})
    _, want = model._precipitation(st, 5.0)
    for n in (mm.PREC, mm.ACCPREC):
        np.testing.assert_array_equal(diags[n][: g.nx, : g.ny], want[n][: g.nx, : g.ny], err_msg=n)
    assert float(want[mm.PREC].max()) > 0.0


def _scalar(v):
    """A stand-in for the grid's DataArray spacings (``grid.dx.to_units('m').values.item()``)."""
    s = types.SimpleNamespace(values=np.float64(v))
    s.to_units = lambda units: s
    return s


def test_coriolis_smagorinsky_and_vertical_advection_array_calls(evolved):
    model, st = evolved
    _, _, gen = modules()
    g, nb = model.g, model.hb.nb
    box = (slice(0, g.nx), slice(0, g.ny), slice(0, g.nz))
    # ---- Coriolis, isentropic/physics/coriolis.py:L139-L164
    co = refload.load("tasmania.isentropic.physics.coriolis").IsentropicConservativeCoriolis
    me = stand_in(model, co._stencil_numpy, {"set_output": gen.set_output_numpy}, _nb=nb, _f=model.f)
    tnd = {mm.SU: model.z(), mm.SV: model.z()}
    co.array_call(me, st, tnd, {}, {mm.SU: True, mm.SV: True})
    want, _ = model._coriolis(st)
    for n in tnd:
        np.testing.assert_array_equal(tnd[n], want[n], err_msg=n)
    assert float(np.abs(want[mm.SV]).max()) > 0.0
    # ---- IsentropicSmagorinsky, isentropic/physics/turbulence.py:L68-L97
    tu2 = refload.load("tasmania.physics.turbulence").Smagorinsky2d
    tui = refload.load("tasmania.isentropic.physics.turbulence").IsentropicSmagorinsky
    me = stand_in(model, tui._stencil_numpy, {"set_output": gen.set_output_numpy, "core": tu2._core_numpy},
                  _nb=max(2, nb), _cs=model.cs)
    me.grid.dx, me.grid.dy = _scalar(g.dx), _scalar(g.dy)
    tnd = {mm.SU: model.z(), mm.SV: model.z()}
    with np.errstate(divide="ignore", invalid="ignore"):
        tui.array_call(me, st, tnd, {}, {mm.SU: True, mm.SV: True})
    want, _ = model._smagorinsky(st)
    for n in tnd:
        np.testing.assert_array_equal(tnd[n], want[n], err_msg=n)
    assert float(np.abs(want[mm.SU]).max()) > 0.0
    # ---- IsentropicVerticalAdvection (moist, w on main levels, third-order upwind),
    # isentropic/physics/vertical_advection.py:L216-L269
    va = refload.load("tasmania.isentropic.physics.vertical_advection").IsentropicVerticalAdvection
    fl = refload.load("tasmania.isentropic.dynamics.subclasses.minimal_vertical_fluxes.third_order_upwind")
    flux_cls = fl.ThirdOrderUpwind
    stencil = refload.numpy_stencil(va._stencil_numpy, {"set_output": gen.set_output_numpy})
    me = types.SimpleNamespace(
        grid=types.SimpleNamespace(nx=g.nx, ny=g.ny, nz=g.nz, dz=_scalar(g.dz)),
        backend_options=types.SimpleNamespace(exec_info=None, validate_args=False),
        _stgz=False, _moist=True,
        _vflux=types.SimpleNamespace(extent=flux_cls.extent,
                                     get_subroutine_definition=lambda name: getattr(flux_cls, name + "_numpy")),
        get_field_storage_shape=lambda name=None: model.shape,
        storage_options=types.SimpleNamespace(dtype=np.float64))
    me._stencil = lambda exec_info=None, validate_args=None, **kw: stencil(me, **kw)
    names = (mm.S, mm.SU, mm.SV, mm.QV, mm.QC, mm.QR)
    tnd = {n: np.full(model.shape, 7.0) for n in names}
    va.array_call(me, st, tnd, {}, {n: True for n in names})
    want, _ = model._vertical_advection(st)
    for n in names:
        np.testing.assert_array_equal(tnd[n], want[n], err_msg=n)
    assert float(np.abs(want[mm.QV][box]).max()) > 0.0


def test_clipping_array_call(evolved):
    """Clipping (physics/microphysics/utils.py:L58-L141): the reference's array_call with its own
    clip_numpy stencil against the b200 mirror through the oracle-backed stub, on species with
    negative values, bit for bit; the inputs stay untouched."""
    import tasmania_b200 as tb
    from tasmania_b200.microphysics import Clipping
    from tests.abi_oracle import OracleStub
    from tests.abi_stub import stubbed_library

    model, st = evolved
    _, ut, _ = modules()
    math = refload.load("tasmania.framework.subclasses.stencil_definitions.math")
    names = (mm.QV, mm.QC, mm.QR)
    rng = np.random.default_rng(8)
    state = {n: st[n] - 0.5 * np.abs(st[n]).mean() * rng.random(st[n].shape) for n in names}
    assert all((v < 0).any() and (v > 0).any() for v in state.values())
    me = stand_in(model, math.clip_numpy, {}, _names=names)
    ref = {n: np.full(model.shape, 7.0) for n in names}
    ut.Clipping.array_call(me, state, ref)
    with stubbed_library(OracleStub) as stub:
        comp = Clipping(model.g, water_species_names=names)
        dev_state = {n: tb.as_storage(v) for n, v in state.items()}
        out = {n: tb.as_storage(np.full(model.shape, 7.0)) for n in names}
        comp.array_call(dev_state, out)
        assert stub.count("tb200_elementwise") == 3
        for n in names:
            np.testing.assert_array_equal(tb.to_numpy(out[n]), ref[n], err_msg=n)
            np.testing.assert_array_equal(tb.to_numpy(dev_state[n]), state[n])
            assert float(ref[n].min()) == 0.0
