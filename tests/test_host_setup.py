# -*- coding: utf-8 -*-
"""CPU checks of the host side: set-up formulas vs the reference fixture, storage wrapper
semantics, registry behaviour, and that the C-ABI library loads and exports every symbol
include/tasmania_b200.h declares (no compute without a GPU)."""
import copy
import os
import re
from datetime import timedelta

import numpy as np
import pytest

import tasmania_b200 as tb
from tasmania_b200 import lib
from tasmania_b200.grid import Grid, Topography, gaussian_profile, isentropic_state_from_brunt_vaisala
from tests import helpers as hp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_setup_matches_reference_fixture():
    fx = hp.load("isen_dry_rk3_5th")
    nx, ny, nz = (int(v) for v in fx["dims"][:3])
    x = np.linspace(-176.0, 176.0, nx)
    y = np.linspace(-176.0, 176.0, ny)
    np.testing.assert_array_equal(gaussian_profile(x, y, 500.0, 50.0, 50.0), fx["topo_steady"])
    grid = Grid((-176.0, 176.0), nx, (-176.0, 176.0), ny, (400.0, 280.0), nz, units_to_m=1e3,
                topography=Topography(fx["topo_steady"], timedelta(seconds=60)))
    np.testing.assert_array_equal(grid.x, fx["x"])
    np.testing.assert_array_equal(grid.z, fx["z"])
    assert grid.dx == float(fx["params"][0]) and grid.dz == float(fx["params"][2])
    st = isentropic_state_from_brunt_vaisala(grid, 22.5, 0.0, 0.015)
    for n in (hp.S, hp.SU, hp.SV, hp.U, hp.V, hp.MTG, hp.P, hp.EXN, hp.H):
        np.testing.assert_array_equal(st[n], fx["init_" + n])


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tasmania_b200.h")).read()
    declared = set(re.findall(r"\b(tb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    l = lib.load()
    for name in declared:
        assert hasattr(l, name), f"{name} declared in the header but not exported"
    assert declared == set(lib.exported_symbols())
    assert l.tb200_version() >= 100


def test_no_cpu_fallback():
    a = np.zeros((4, 4, 2))
    with pytest.raises(tb.B200Error):
        tb.compile_stencil("copy")(src=a, dst=a, origin=(0, 0, 0), domain=(4, 4, 2))
    with pytest.raises(tb.FactoryRegistryError):
        tb.compile_stencil("copy", backend="numpy")
    with pytest.raises(tb.FactoryRegistryError):
        tb.compile_stencil("not_a_stencil")


def test_storage_wrapper_semantics():
    a = tb.zeros((5, 4, 3), device="cpu")
    assert a.shape == (5, 4, 3) and a.dtype == np.float64
    assert a.strides == (8, 16 * 8, 16 * 4 * 8)  # i fastest, rows padded to 16 doubles
    a[1:3, :, 0] = 2.0
    a[:, :, 1] = np.arange(20.0).reshape(5, 4)
    a[...] = a  # self assignment is a no-op
    b = tb.as_storage(np.arange(3.0)[None, None, :], device="cpu")
    a[:2, :2, :] = b  # broadcast
    ref = np.zeros((5, 4, 3))
    ref[1:3, :, 0] = 2.0
    ref[:, :, 1] = np.arange(20.0).reshape(5, 4)
    ref[:2, :2, :] = np.arange(3.0)
    np.testing.assert_array_equal(tb.to_numpy(a), ref)
    v = a[:, :, 1:2]
    assert v.shape == (5, 4, 1) and v.strides == a.strides
    assert a[4, 3, 1] == 19.0
    c = copy.deepcopy(a)
    c[0, 0, 0] = -1.0
    assert a[0, 0, 0] != -1.0
    np.testing.assert_array_equal(np.asarray(c)[1:], ref[1:])


def _fma(a, b, c):
    """Correctly rounded a * b + c (exact rational arithmetic, then one rounding)."""
    from fractions import Fraction

    return float(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def test_constant_division_is_correctly_rounded():
    """csrc/common.cuh:CDiv -- q0 = a * rc; r = fma(-c, q0, a); q = fma(r, rc, q0) must return
    RN(a / c), i.e. numpy's a / c bit for bit, for the constants the stencils divide by."""
    rng = np.random.default_rng(7)
    for c in (12.0, 60.0, 2200.0, 4400.0, 1e5, 3437.5, 2.0 * 3437.5, 1.0, 1000.0 / 3.0):
        rc = 1.0 / c
        vals = np.concatenate([rng.uniform(-1e6, 1e6, 300), rng.uniform(-1e-3, 1e-3, 100),
                               rng.standard_normal(100) * 1e9])
        for a in vals:
            q0 = float(a) * rc
            r = _fma(-c, q0, a)
            q = _fma(r, rc, q0)
            assert q == float(a) / c, (a, c)


def _pow_pos_restated(x, kappa):
    """csrc/stencil_math.cuh:pow_pos, operation by operation (the reciprocal seed is replaced by
    a perturbed 1/d: the Newton step and the residual correction must absorb it)."""
    import struct
    from fractions import Fraction

    log_c = [float(Fraction(2, 2 * k + 3)) for k in range(9)]
    from decimal import Decimal, getcontext
    import math

    getcontext().prec = 60
    ln2 = Decimal(2).ln()
    exp_c = [float(ln2 ** k / Decimal(math.factorial(k))) for k in range(14)]
    l2e = float(Decimal(1) / ln2)
    bits = struct.unpack("<q", struct.pack("<d", x))[0]
    hi = bits >> 32
    e = (hi - 0x3FE6A09F) >> 20
    m = struct.unpack("<d", struct.pack("<q", bits - (e << 52)))[0]
    assert 0.70710 < m < 1.41422
    f, d = m - 1.0, m + 1.0
    r = (1.0 / d) * (1.0 + 2.0 ** -21)  # MUFU.RCP64H-grade seed
    t = _fma(-d, r, 1.0)
    t = _fma(t, t, t)
    r = _fma(r, t, r)
    s = f * r
    s = _fma(_fma(-d, s, f), r, s)
    z = s * s
    p = log_c[8]
    for n in range(7, -1, -1):
        p = _fma(p, z, log_c[n])
    lm = _fma(s * z, p, 2.0 * s)
    ef = float(e)
    l2 = _fma(lm, l2e, ef)
    l2_lo = _fma(lm, l2e, ef - l2)
    y = kappa * l2
    y_lo = _fma(kappa, l2_lo, _fma(kappa, l2, -y))
    shifter = 6755399441055744.0
    ts = y + shifter
    n = int(ts - shifter)
    q = (y - (ts - shifter)) + y_lo
    w = exp_c[13]
    for k in range(12, -1, -1):
        w = _fma(w, q, exp_c[k])
    return math.ldexp(w, n)


def test_exner_power_restatement_accuracy():
    """The domain-specific x**kappa of the column scans: relative error against 60-digit
    arithmetic <= 1.7e-16 (glibc's pow, which the reference calls, reaches 1.3e-16 on the same
    samples), over the pressures met in the model and far outside them."""
    from decimal import Decimal, getcontext

    getcontext().prec = 60
    rng = np.random.default_rng(11)
    kappa = 287.05 / 1004.0
    worst = 0.0
    for lo, hi in ((0.005, 0.1), (0.1, 0.6), (0.6, 0.72), (0.72, 1.3), (1.3, 4.0),
                   (1e-300, 1e-290), (1e200, 1e300)):
        for x in rng.uniform(lo, hi, 120):
            got = _pow_pos_restated(float(x), kappa)
            ref = (Decimal(float(x)).ln() * Decimal(kappa)).exp()
            worst = max(worst, float(abs(Decimal(got) - ref) / ref))
    assert worst <= 1.7e-16, worst


def test_identity_boundary_mirror_touches_nothing():
    """identity.py:L30-L79: numerical grid == physical grid, every operation a no-op (and no
    library call: the recording stub stays empty)."""
    import numpy as np

    from tasmania_b200.boundary import HorizontalBoundary
    from tests.abi_stub import stubbed_library

    with stubbed_library() as stub:
        hb = HorizontalBoundary.factory("identity", 9, 7, 3, 2)
        assert (hb.ni, hb.nj, hb.type) == (9, 7, "identity")
        a = np.arange(9 * 7 * 3, dtype=float).reshape(9, 7, 3)
        b = a.copy()
        assert hb.get_numerical_field(a) is a and hb.get_physical_field(a) is a
        hb.enforce_field(a, "air_isentropic_density")
        hb.set_outermost_layers_x(a)
        hb.set_outermost_layers_y(a)
        hb.enforce_raw({"air_isentropic_density": a})
        assert (a == b).all() and stub.calls == []


def test_stage_scratch_keeps_plain_storages_without_a_device(monkeypatch):
    """storage.stage_scratch hands out library-context fields on the current CUDA device only; the
    CPU test double of the library (DEFAULT_DEVICE_OVERRIDE) gets ordinary storages and no context."""
    from tasmania_b200 import storage

    monkeypatch.setattr(storage, "DEFAULT_DEVICE_OVERRIDE", "cpu")
    ctx, fields = storage.stage_scratch((7, 5, 3), 3)
    assert ctx is None and len(fields) == 3
    for f in fields:
        assert isinstance(f, storage.B200Array) and f.shape == (7, 5, 3) and f.strides == (8, 16 * 8, 16 * 5 * 8)
        assert float(np.abs(tb.to_numpy(f)).max()) == 0.0
    assert len({f.t.data_ptr() for f in fields}) == 3


def test_dycore_mirror_accepts_the_reference_smoothing_arguments(monkeypatch):
    """A reference call site passes smooth* keywords (dycore.py:L81-L92); the reference keeps the
    flags and never smooths inside a stage, and so does the mirror.  Unknown keywords still raise."""
    from tasmania_b200 import storage
    from tasmania_b200.boundary import Relaxed
    from tasmania_b200.isentropic import IsentropicDynamicalCore

    monkeypatch.setattr(storage, "DEFAULT_DEVICE_OVERRIDE", "cpu")
    x = np.linspace(-176.0, 176.0, 19)
    grid = Grid((-176.0, 176.0), 19, (-176.0, 176.0), 17, (400.0, 280.0), 6, units_to_m=1e3,
                topography=Topography(gaussian_profile(x, np.linspace(-176.0, 176.0, 17), 500.0, 50.0, 50.0),
                                      timedelta(seconds=30)))
    kw = dict(time_integration_scheme="rk3ws_si", horizontal_flux_scheme="fifth_order_upwind",
              time_integration_properties={"pt": 100.0, "eps": 0.5}, damp_depth=2, fused=False)
    dyc = IsentropicDynamicalCore(grid, Relaxed(19, 17, 6, 3, nr=6), smooth=False, smooth_type="second_order",
                                  smooth_coeff=0.2, smooth_moist=False, smooth_moist_damp_depth=0, **kw)
    assert dyc._smooth is False and dyc._smooth_moist is False
    with pytest.raises(TypeError):
        IsentropicDynamicalCore(grid, Relaxed(19, 17, 6, 3, nr=6), smoooth=False, **kw)


def test_stencil_factory_compiler_accessors():
    """framework/stencil.py:L286-L297, L356-L427: get_stencil_compiler / get_subroutine_compiler /
    compile_subroutine exist on every mirror (one compiler per kind: there is one backend)."""
    from tasmania_b200 import framework as fw

    f = fw.StencilFactory()
    assert f.get_stencil_compiler("b200") is fw.compiler_b200
    assert f.get_subroutine_compiler("b200", "flux_dry") is fw.subroutine_compiler_b200
    with pytest.raises(tb.FactoryRegistryError):
        f.get_stencil_compiler("numpy")
    desc = f.get_subroutine_definition("set_output")
    assert f.compile_subroutine("set_output") is desc


def test_physical_constants_by_the_reference_names(monkeypatch):
    """IsentropicDiagnostics takes its constants under the reference's names too
    (isentropic/dynamics/diagnostics.py:L56-L63) and reports them through raw_physical_constants."""
    from tasmania_b200 import storage
    from tasmania_b200.isentropic import IsentropicDiagnostics

    monkeypatch.setattr(storage, "DEFAULT_DEVICE_OVERRIDE", "cpu")
    x, y = np.linspace(-1.0, 1.0, 9), np.linspace(-1.0, 1.0, 7)
    grid = Grid((-1.0, 1.0), 9, (-1.0, 1.0), 7, (400.0, 280.0), 4, units_to_m=1e3,
                topography=Topography(gaussian_profile(x, y, 500.0, 0.5, 0.5), timedelta(seconds=0)))
    d = IsentropicDiagnostics(grid)
    assert d.raw_physical_constants == d.default_physical_constants
    assert d.raw_physical_constants["gravitational_acceleration"] == 9.80665
    d = IsentropicDiagnostics(grid, {"gravitational_acceleration": 9.81, "cp": 1005.0})
    assert d.rpc["g"] == 9.81 and d.rpc["cp"] == 1005.0 and d.rpc["rd"] == 287.05
    assert d.raw_physical_constants["specific_heat_of_dry_air_at_constant_pressure"] == 1005.0
    assert d.backend_options.externals["g"] == 9.81
