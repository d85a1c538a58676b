# -*- coding: utf-8 -*-
"""Property tests in the reference's own style (``@given(data=...)`` drawing
random grids of 1-20 points per axis, nb <= 4 and fields in the ranges of the reference's
tests/conf.py:L99-L175 -- s in [10, 1000], u, v in [-50, 50], q in [0, 5]).  Where the reference
compares its components with inline numpy restatements at rtol = 1e-5, here the oracle is compared
with the reference's own numpy stencils and boundary classes run in place, bit for bit.  Skipped
where /root/reference is absent."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import boundary as ob
from oracle import isentropic as oi
from oracle.fluxes import EXTENT
from tests.golden import refload
from tests.test_boundaries_reference import FIELDS, _domain
from tests.test_stencils_reference import CLASSES, _reference_stencils

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference tree not mounted")

SETTINGS = dict(deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)


@settings(max_examples=30, **SETTINGS)
@given(data=st.data())
def test_k1_k2_on_random_grids(data):
    """step_forward_euler / step_forward_euler_momentum (prognostics/utils.py:L43-L204) with a random
    scheme, grid, compute box inside the storage, time step, spacing and tendencies."""
    scheme = data.draw(st.sampled_from(sorted(CLASSES)), label="scheme")
    moist = data.draw(st.booleans(), label="moist")
    tnd = data.draw(st.booleans(), label="tendencies")
    e = EXTENT[scheme]
    nx = data.draw(st.integers(2 * e + 1, 20), label="nx")
    ny = data.draw(st.integers(2 * e + 1, 20), label="ny")
    nz = data.draw(st.integers(1, 6), label="nz")
    i0 = data.draw(st.integers(e, nx - e - 1), label="i0")
    j0 = data.draw(st.integers(e, ny - e - 1), label="j0")
    k0 = data.draw(st.integers(0, nz - 1), label="k0")
    origin = (i0, j0, k0)
    domain = (data.draw(st.integers(1, nx - e - i0)), data.draw(st.integers(1, ny - e - j0)),
              data.draw(st.integers(1, nz - k0)))
    dt = data.draw(st.floats(0.1, 60.0), label="dt")
    dx = data.draw(st.floats(50.0, 5000.0), label="dx")
    dy = data.draw(st.floats(50.0, 5000.0), label="dy")
    eps = data.draw(st.floats(0.0, 1.0), label="eps")
    rng = np.random.default_rng(data.draw(st.integers(0, 2**31), label="seed"))
    shape = (nx + 1, ny + 1, nz + 1)

    def field(lo, hi):
        return rng.uniform(lo, hi, size=shape)

    k1, k2 = _reference_stencils(scheme, moist)
    s_now, s_int, s_new_in = field(10, 1000), field(10, 1000), field(10, 1000)
    u, v = field(-50, 50), field(-50, 50)
    su_now, su_int, sv_now, sv_int = (field(-5e4, 5e4) for _ in range(4))
    mtg_now, mtg_new = field(2.9e5, 3.1e5), field(2.9e5, 3.1e5)
    s_tnd, su_tnd, sv_tnd = field(-1, 1), field(-10, 10), field(-10, 10)
    sq_now, sq_int = [field(0, 5000) for _ in range(3)], [field(0, 5000) for _ in range(3)]
    q_tnd = [field(-1e-3, 1e-3) for _ in range(3)]
    ref_s = field(0, 1)
    ora_s = ref_s.copy()
    ref_q = [field(0, 1) for _ in range(3)]
    ora_q = [q.copy() for q in ref_q]
    kw_ref, kw_ora = {}, {}
    if moist:
        for i, t in enumerate(("qv", "qc", "qr")):
            kw_ref.update({f"s{t}_now": sq_now[i], f"s{t}_int": sq_int[i], f"s{t}_new": ref_q[i]})
            if tnd:
                kw_ref[f"{t}_tnd"] = q_tnd[i]
        kw_ora = dict(moist=True, sq_now=sq_now, sq_int=sq_int, sq_new=ora_q,
                      q_tnd=q_tnd if tnd else (None, None, None))
    k1(s_now=s_now, s_int=s_int, s_new=ref_s, u_int=u, v_int=v, su_int=su_int, sv_int=sv_int,
       s_tnd=s_tnd if tnd else None, dt=dt, dx=dx, dy=dy, origin=origin, domain=domain, **kw_ref)
    oi.step_forward_euler(scheme, s_now, s_int, ora_s, u, v, dt=dt, dx=dx, dy=dy, origin=origin,
                          domain=domain, s_tnd=s_tnd if tnd else None, **kw_ora)
    np.testing.assert_array_equal(ora_s, ref_s)
    for a, b in zip(ora_q, ref_q):
        np.testing.assert_array_equal(a, b)
    if moist:
        return
    ref_su, ref_sv = field(0, 1), field(0, 1)
    ora_su, ora_sv = ref_su.copy(), ref_sv.copy()
    k2(s_now=s_now, s_int=s_int, s_new=s_new_in, u_int=u, v_int=v, su_now=su_now, su_int=su_int,
       su_new=ref_su, sv_now=sv_now, sv_int=sv_int, sv_new=ref_sv, mtg_now=mtg_now, mtg_new=mtg_new,
       su_tnd=su_tnd if tnd else None, sv_tnd=sv_tnd if tnd else None, dt=dt, dx=dx, dy=dy, eps=eps,
       origin=origin, domain=domain)
    oi.step_forward_euler_momentum(
        scheme, s_now, s_new_in, u, v, su_now, su_int, ora_su, sv_now, sv_int, ora_sv, mtg_now, mtg_new,
        dt=dt, dx=dx, dy=dy, eps=eps, origin=origin, domain=domain,
        su_tnd=su_tnd if tnd else None, sv_tnd=sv_tnd if tnd else None)
    np.testing.assert_array_equal(ora_su, ref_su)
    np.testing.assert_array_equal(ora_sv, ref_sv)


@settings(max_examples=20, **SETTINGS)
@given(data=st.data())
def test_periodic_and_relaxed_boundaries_on_random_grids(data):
    """Periodic (periodic.py:L32-L122) and Relaxed (relaxed.py:L119-L247) on random grids, boundary
    depths and (staggered) fields: numerical <-> physical fields, enforce_field, outermost layers."""
    kind = data.draw(st.sampled_from(("periodic", "relaxed")), label="kind")
    nb = data.draw(st.integers(1, 4), label="nb")
    nr = data.draw(st.integers(nb, 8), label="nr")
    lo = 2 * (nb if kind == "periodic" else nr)
    nx = data.draw(st.integers(max(lo, 2), 20), label="nx")
    ny = data.draw(st.integers(max(lo, 2), 20), label="ny")
    nz = data.draw(st.integers(1, 5), label="nz")
    rng = np.random.default_rng(data.draw(st.integers(0, 2**31), label="seed"))
    if kind == "periodic":
        hb = _domain(nx, ny, nz, "periodic", nb).horizontal_boundary
        ohb = ob.Periodic(nx, ny, nz, nb)
        for name in FIELDS:
            mx = nx + ("at_u_locations" in name)
            my = ny + ("at_v_locations" in name)
            phys = rng.standard_normal((mx, my, nz))
            num = np.asarray(hb.get_numerical_field(phys.copy(), field_name=name))
            np.testing.assert_array_equal(ohb.get_numerical_field(phys.copy(), name), num, err_msg=name)
            a = rng.standard_normal(num.shape)
            b = a.copy()
            hb.enforce_field(a, field_name=name)
            ohb.enforce_field(b, name)
            np.testing.assert_array_equal(b, a, err_msg=name)
        return
    hb = _domain(nx, ny, nz, "relaxed", nb, nr=nr).horizontal_boundary
    ohb = ob.Relaxed(nx, ny, nz, nb, nr)
    np.testing.assert_array_equal(ohb.gamma[: nx + 1, : ny + 1, 0], np.asarray(hb._gamma)[: nx + 1, : ny + 1, 0])
    shape = (nx + 1, ny + 1, nz + 1)
    ref_state = {n: rng.standard_normal(shape) for n in FIELDS}
    hb.reference_state = {n: refload.DataArray(v.copy(), attrs={"units": "1"}) for n, v in ref_state.items()}
    ohb.reference_state = {n: v.copy() for n, v in ref_state.items()}
    for name in FIELDS:
        a = rng.standard_normal(shape)
        b = a.copy()
        hb.enforce_field(a, field_name=name, field_units="1")
        ohb.enforce_field(b, name)
        np.testing.assert_array_equal(b, a, err_msg=name)
    for name, ref_fn, ora_fn in (
            ("x_velocity_at_u_locations", hb.set_outermost_layers_x, ohb.set_outermost_layers_x),
            ("y_velocity_at_v_locations", hb.set_outermost_layers_y, ohb.set_outermost_layers_y)):
        a = rng.standard_normal(shape)
        b = a.copy()
        ref_fn(a, field_name=name, field_units="1")
        ora_fn(b, name)
        np.testing.assert_array_equal(b, a, err_msg=name)


@settings(max_examples=15, **SETTINGS)
@given(shape=st.tuples(st.integers(3, 20), st.integers(3, 20), st.integers(1, 5)))
def test_diffusers_and_smoothers_on_random_shapes(shape):
    """HorizontalDiffusion / HorizontalSmoothing (2-D, 1DX, 1DY; every order the shape admits)."""
    from tests import test_stencils_reference as ts

    ts.test_diffusers_and_smoothers_equal_reference_on_minimal_and_ragged_grids(shape)


@settings(max_examples=15, **SETTINGS)
@given(dims=st.tuples(st.integers(1, 12), st.integers(1, 12), st.integers(1, 72)))
def test_k3_column_scans_on_random_columns(dims):
    """IsentropicDiagnostics' four numpy stencils on random column counts and depths."""
    from tests import test_stencils_reference as ts

    ts.test_k3_column_scans_equal_reference_from_one_level_to_deep_columns(dims)


@settings(max_examples=12, **SETTINGS)
@given(data=st.data())
def test_b200_mirror_components_on_random_grids(data):
    """The reference's tests run every component on every backend; here the b200 mirrors
    (HorizontalDiffusion, HorizontalSmoothing, Periodic, Relaxed: host marshalling + the C-ABI call,
    carried out by the oracle-backed stub) against the reference's own classes on backend numpy, on
    random grids -- bit for bit."""
    import tasmania_b200 as tb
    from tasmania_b200 import boundary as tbb
    from tasmania_b200.dwarfs import HorizontalDiffusion, HorizontalSmoothing
    from tests.abi_oracle import OracleStub
    from tests.abi_stub import stubbed_library

    refload.install_framework()
    opts = refload.load("tasmania.framework.options")
    for name in ("second_order", "fourth_order"):
        refload.load("tasmania.dwarfs.subclasses.horizontal_diffusers." + name)
    for name in ("first_order", "second_order", "third_order"):
        refload.load("tasmania.dwarfs.subclasses.horizontal_smoothers." + name)
    hd = refload.load("tasmania.dwarfs.horizontal_diffusion")
    hsm = refload.load("tasmania.dwarfs.horizontal_smoothing")
    kw = dict(backend="numpy", backend_options=opts.BackendOptions(), storage_options=opts.StorageOptions())

    nx = data.draw(st.integers(7, 20), label="nx")
    ny = data.draw(st.integers(7, 20), label="ny")
    nz = data.draw(st.integers(1, 5), label="nz")
    depth = data.draw(st.integers(0, nz), label="damp depth")
    nb = data.draw(st.integers(1, 3), label="nb")
    rng = np.random.default_rng(data.draw(st.integers(0, 2**31), label="seed"))
    shape = (nx, ny, nz)
    phi, base = rng.standard_normal(shape), rng.standard_normal(shape)
    with stubbed_library(OracleStub):
        for name in ("second_order", "fourth_order"):
            ref_obj = hd.HorizontalDiffusion.factory(name, shape, 0.7, 1.3, 0.5, 1.0, depth, **kw)
            obj = HorizontalDiffusion.factory(name, shape, 0.7, 1.3, 0.5, 1.0, depth)
            for overwrite in (True, False):
                ref = base.copy()
                ref_obj(phi, ref, overwrite_output=overwrite)
                out = tb.as_storage(base)
                obj(tb.as_storage(phi), out, overwrite_output=overwrite)
                np.testing.assert_array_equal(tb.to_numpy(out), ref, err_msg=f"diffusion {name} {overwrite}")
        for name in ("first_order", "second_order", "third_order"):
            ref_obj = hsm.HorizontalSmoothing.factory(name, shape, 0.03, 0.24, depth, **kw)
            obj = HorizontalSmoothing.factory(name, shape, 0.03, 0.24, depth)
            ref = base.copy()
            ref_obj(phi, ref)
            out = tb.as_storage(base)
            obj(tb.as_storage(phi), out)
            np.testing.assert_array_equal(tb.to_numpy(out), ref, err_msg=f"smoothing {name}")
        # periodic wrap of unstaggered and staggered fields on the numerical grid
        px, py = max(nx - 2 * nb, 2 * nb), max(ny - 2 * nb, 2 * nb)
        hb_ref = _domain(px, py, nz, "periodic", nb).horizontal_boundary
        hb = tbb.Periodic(px, py, nz, nb)
        for name in FIELDS:
            a = rng.standard_normal((px + 2 * nb + 1, py + 2 * nb + 1, nz + 1))
            b = tb.as_storage(a)
            hb_ref.enforce_field(a, field_name=name)
            hb.enforce_field(b, field_name=name)
            np.testing.assert_array_equal(tb.to_numpy(b), a, err_msg="periodic " + name)


@settings(max_examples=40, **SETTINGS)
@given(data=st.data())
def test_grid_component_shape_helpers(data):
    """framework.GridComponent (the shape helpers every grid-based mirror inherits) against the
    reference's GridComponent methods run in place (framework/base_components.py:L55-L135)."""
    import types

    from tasmania_b200.framework import GridComponent

    refload.install_framework()
    ref_cls = refload.load("tasmania.framework.base_components").GridComponent
    grid = types.SimpleNamespace(nx=data.draw(st.integers(1, 20)), ny=data.draw(st.integers(1, 20)),
                                 nz=data.draw(st.integers(1, 20)))
    ref_self = types.SimpleNamespace(grid=grid, get_field_grid_shape=None, get_shape=ref_cls.get_shape)
    ref_self.get_field_grid_shape = lambda name: ref_cls.get_field_grid_shape(ref_self, name)
    mine = GridComponent()
    mine.grid = grid
    stem = data.draw(st.sampled_from(("air_density", "x_velocity", "precipitation", "height")))
    stag = data.draw(st.sampled_from(("", "_at_u_locations", "_at_v_locations", "_at_uv_locations")))
    lev = data.draw(st.sampled_from(("", "_on_interface_levels", "_at_surface_level")))
    name = stem + stag + lev
    assert tuple(mine.get_field_grid_shape(name)) == tuple(ref_cls.get_field_grid_shape(ref_self, name))
    triple = st.tuples(st.integers(1, 24), st.integers(1, 24), st.integers(1, 24))
    shape = data.draw(st.one_of(st.none(), triple), label="shape")
    lo = data.draw(triple, label="min")
    hi = data.draw(st.one_of(st.none(), triple.map(lambda t: tuple(max(a, b) for a, b in zip(t, lo)))), label="max")
    assert list(mine.get_shape(shape, lo, hi)) == list(ref_cls.get_shape(shape, lo, hi))
    assert list(mine.get_field_storage_shape(name, shape)) == list(
        ref_cls.get_field_storage_shape(ref_self, name, shape))
    assert list(mine.get_storage_shape(shape, None, hi and tuple(max(a, b) for a, b in zip(hi, (grid.nx, grid.ny, grid.nz))))) == list(
        ref_cls.get_storage_shape(ref_self, shape, None, hi and tuple(max(a, b) for a, b in zip(hi, (grid.nx, grid.ny, grid.nz)))))
