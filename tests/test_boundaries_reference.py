# -*- coding: utf-8 -*-
"""The lateral boundaries of the oracle against the REFERENCE's own classes run in place (skipped
where /root/reference is absent): ``Periodic`` (periodic.py:L32-L122: numerical <-> physical fields,
enforce_field, outermost layers; unstaggered and staggered fields) and ``Relaxed``
(relaxed.py:L119-L247: coefficient matrix, enforce_field, outermost layers) -- bit for bit.  The GPU
kernels are held to the oracle in tests/test_gpu_stencils.py and tests/test_gpu_relax_frame.py."""
import numpy as np
import pytest

from oracle import boundary as ob
from tests.golden import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference tree not mounted")

FIELDS = ("air_isentropic_density", "x_velocity_at_u_locations", "y_velocity_at_v_locations")


def _domain(nx, ny, nz, kind, nb, **kw):
    from tests.golden import generate_golden as gg

    refload.install_framework()
    return gg._make_domain(nx, ny, nz, kind, nb, kw, topo=False)


@pytest.mark.parametrize("nb", [1, 2, 3])
def test_periodic_boundary_equals_reference(nb):
    nx, ny, nz = 13, 11, 4
    hb = _domain(nx, ny, nz, "periodic", nb).horizontal_boundary
    ohb = ob.Periodic(nx, ny, nz, nb)
    assert (hb.ni, hb.nj) == (ohb.ni, ohb.nj)
    rng = np.random.default_rng(nb)
    for name in FIELDS:
        mx = nx + ("at_u_locations" in name)
        my = ny + ("at_v_locations" in name)
        phys = rng.standard_normal((mx, my, nz))
        num = np.asarray(hb.get_numerical_field(phys.copy(), field_name=name))
        np.testing.assert_array_equal(ohb.get_numerical_field(phys.copy(), name), num, err_msg=name)
        np.testing.assert_array_equal(ohb.get_physical_field(num, name),
                                      np.asarray(hb.get_physical_field(num, field_name=name)))
        a = rng.standard_normal(num.shape)
        b = a.copy()
        hb.enforce_field(a, field_name=name)
        ohb.enforce_field(b, name)
        np.testing.assert_array_equal(b, a, err_msg=name)
    for axis, ref_fn, ora_fn in ((0, hb.set_outermost_layers_x, ohb.set_outermost_layers_x),
                                 (1, hb.set_outermost_layers_y, ohb.set_outermost_layers_y)):
        a = rng.standard_normal((nx + 2 * nb + 1, ny + 2 * nb + 1, nz))
        b = a.copy()
        ref_fn(a, field_name="x_velocity_at_u_locations" if axis == 0 else "y_velocity_at_v_locations")
        ora_fn(b)
        np.testing.assert_array_equal(b, a)


@pytest.mark.parametrize("nb,nr", [(3, 6), (2, 8), (1, 4)])
def test_relaxed_boundary_equals_reference(nb, nr):
    nx, ny, nz = 21, 19, 5
    hb = _domain(nx, ny, nz, "relaxed", nb, nr=nr).horizontal_boundary
    ohb = ob.Relaxed(nx, ny, nz, nb, nr)
    np.testing.assert_array_equal(ohb.gamma[: nx + 1, : ny + 1, 0], np.asarray(hb._gamma)[: nx + 1, : ny + 1, 0])
    rng = np.random.default_rng(nr)
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


@pytest.mark.parametrize("kind,kw,nb", [("relaxed", {"nr": 5}, 2), ("relaxed", {"nr": 3}, 3),
                                        ("periodic", {}, 2), ("periodic", {}, 1)])
@pytest.mark.parametrize("nx,ny", [(17, 1), (1, 15)])
def test_one_dimensional_boundary_mirrors_equal_reference(kind, kw, nb, nx, ny):
    """The b200 mirrors of Relaxed1DX / 1DY and Periodic1DX / 1DY (tasmania_b200/boundary.py), run
    numerically through the oracle-backed ABI stub, against the reference's own classes (picked by
    its factory on ny == 1 / nx == 1 grids): numerical <-> physical fields, enforce_field on
    unstaggered / staggered / interface-level fields, outermost layers -- bit for bit."""
    import tasmania_b200 as tb
    from tasmania_b200.boundary import HorizontalBoundary
    from tests.abi_oracle import OracleStub
    from tests.abi_stub import stubbed_library

    nz = 4
    names = FIELDS + ("air_pressure_on_interface_levels",)
    hb = _domain(nx, ny, nz, kind, nb, **kw).horizontal_boundary
    assert type(hb).__name__.endswith("1DX" if ny == 1 else "1DY")
    rng = np.random.default_rng(nx + nb)
    shape = (hb.ni + 1, hb.nj + 1, nz + 1)
    fields = {n: rng.standard_normal(shape) for n in names}
    refs = {n: rng.standard_normal(shape) for n in names}
    phys = rng.standard_normal((nx, ny, nz))
    hb.reference_state = {n: refload.DataArray(v.copy(), attrs={"units": "1"}) for n, v in refs.items()}
    want = {}
    for n in names:
        f = fields[n].copy()
        hb.enforce_field(f, field_name=n, field_units="1")
        hb.set_outermost_layers_x(f, field_name=n, field_units="1")
        hb.set_outermost_layers_y(f, field_name=n, field_units="1")
        want[n] = f
    want_num = np.asarray(hb.get_numerical_field(phys.copy(), field_name=FIELDS[0]))
    with stubbed_library(OracleStub):
        mhb = HorizontalBoundary.factory(kind, nx, ny, nz, nb, **kw)
        assert type(mhb).__name__ == type(hb).__name__ and (mhb.ni, mhb.nj) == (hb.ni, hb.nj)
        mhb.reference_state = {n: tb.as_storage(v) for n, v in refs.items()}
        for n in names:
            f = tb.as_storage(fields[n])
            mhb.enforce_field(f, field_name=n)
            mhb.set_outermost_layers_x(f, field_name=n)
            mhb.set_outermost_layers_y(f, field_name=n)
            np.testing.assert_array_equal(tb.to_numpy(f), want[n], err_msg=n)
        num = mhb.get_numerical_field(tb.as_storage(phys), field_name=FIELDS[0])
        np.testing.assert_array_equal(tb.to_numpy(num), want_num)
        np.testing.assert_array_equal(
            tb.to_numpy(mhb.get_physical_field(num)),
            np.asarray(hb.get_physical_field(want_num, field_name=FIELDS[0])))
