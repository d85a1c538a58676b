# -*- coding: utf-8 -*-
"""The prognostic-step stencils K1 / K2 of the oracle against the REFERENCE's own numpy definitions
run in place (``step_forward_euler_numpy``, ``step_forward_euler_momentum_numpy``:
isentropic/dynamics/subclasses/prognostics/utils.py:L43-L204 with the four minimal flux classes) on
sizes the committed fixture (one 17x15x6 grid) does not cover: the smallest grid every scheme
accepts (nx = ny = 2 extent + 1: a single computed point), a single level, ragged extents, and a
compute box narrower than the storage -- bit for bit.  Skipped where /root/reference is absent."""
import numpy as np
import pytest

from oracle import isentropic as oi
from oracle.fluxes import EXTENT
from tests.golden import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference tree not mounted")

CLASSES = {"upwind": "Upwind", "centered": "Centered", "third_order_upwind": "ThirdOrderUpwind",
           "fifth_order_upwind": "FifthOrderUpwind"}


def _reference_stencils(scheme, moist):
    refload.install_framework()
    utils = refload.load("tasmania.isentropic.dynamics.subclasses.prognostics.utils")
    mod = refload.load("tasmania.isentropic.dynamics.subclasses.minimal_horizontal_fluxes." + scheme)
    hflux = getattr(mod, CLASSES[scheme])(backend="numpy")
    if scheme == "centered":  # the numpy Centered flux looks its helpers up as globals
        full = refload.load("tasmania.isentropic.dynamics.subclasses.horizontal_fluxes.centered")
        mod.get_centered_flux_x = full.Centered.get_centered_flux_x_numpy
        mod.get_centered_flux_y = full.Centered.get_centered_flux_y_numpy
    ext = dict(hflux.externals or {})
    ext.update(extent=hflux.extent, moist=moist,
               flux_dry=hflux.get_subroutine_definition("flux_dry"),
               flux_moist=hflux.get_subroutine_definition("flux_moist"))
    assert hflux.extent == EXTENT[scheme]
    return (refload.numpy_stencil(utils.step_forward_euler_numpy, ext),
            refload.numpy_stencil(utils.step_forward_euler_momentum_numpy, ext))


def _cases(e):
    n = 2 * e + 1
    # (storage-defining nx, ny, nz), origin, domain
    yield (n, n, 1), (e, e, 0), (1, 1, 1)                      # one computed point, one level
    yield (n + 1, n + 4, 3), (e, e, 0), (1, 4, 3)              # one column of points
    yield (23, 9, 2), (e, e, 0), (23 - 2 * e, 9 - 2 * e, 2)    # ragged
    yield (19, 21, 4), (e + 2, e + 1, 1), (5, 7, 2)            # box strictly inside the storage


@pytest.mark.parametrize("scheme", list(CLASSES))
@pytest.mark.parametrize("moist", (False, True))
def test_k1_k2_equal_reference_on_minimal_and_ragged_grids(scheme, moist):
    k1, k2 = _reference_stencils(scheme, moist)
    e = EXTENT[scheme]
    for n, ((nx, ny, nz), origin, domain) in enumerate(_cases(e)):
        rng = np.random.default_rng(1000 * e + 10 * n + moist)
        shape = (nx + 1, ny + 1, nz + 1)

        def field(lo, hi):
            return rng.uniform(lo, hi, size=shape)

        s_now, s_int, s_new_in = field(10, 1000), field(10, 1000), field(10, 1000)
        u, v = field(-50, 50), field(-50, 50)
        su_now, su_int, sv_now, sv_int = (field(-5e3, 5e3) for _ in range(4))
        mtg_now, mtg_new = field(2.9e5, 3.1e5), field(2.9e5, 3.1e5)
        s_tnd, su_tnd, sv_tnd = field(-1, 1), field(-10, 10), field(-10, 10)
        sq_now, sq_int = [field(0, 5) for _ in range(3)], [field(0, 5) for _ in range(3)]
        q_tnd = [field(-1e-3, 1e-3) for _ in range(3)]
        dt, dx, dy, eps = 1.5 + n, 1100.0, 900.0, 0.3
        for tnd in (False, True):
            ref_s, ora_s = field(0, 1), None
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
            np.testing.assert_array_equal(ora_s, ref_s, err_msg=f"s case {n} tnd {tnd}")
            for a, b in zip(ora_q, ref_q):  # untouched outside the box, equal inside
                np.testing.assert_array_equal(a, b, err_msg=f"sq case {n} tnd {tnd}")
            if moist:
                continue
            ref_su, ref_sv = field(0, 1), field(0, 1)
            ora_su, ora_sv = ref_su.copy(), ref_sv.copy()
            k2(s_now=s_now, s_int=s_int, s_new=s_new_in, u_int=u, v_int=v, su_now=su_now, su_int=su_int,
               su_new=ref_su, sv_now=sv_now, sv_int=sv_int, sv_new=ref_sv, mtg_now=mtg_now,
               mtg_new=mtg_new, su_tnd=su_tnd if tnd else None, sv_tnd=sv_tnd if tnd else None,
               dt=dt, dx=dx, dy=dy, eps=eps, origin=origin, domain=domain)
            oi.step_forward_euler_momentum(
                scheme, s_now, s_new_in, u, v, su_now, su_int, ora_su, sv_now, sv_int, ora_sv, mtg_now,
                mtg_new, dt=dt, dx=dx, dy=dy, eps=eps, origin=origin, domain=domain,
                su_tnd=su_tnd if tnd else None, sv_tnd=sv_tnd if tnd else None)
            np.testing.assert_array_equal(ora_su, ref_su, err_msg=f"su case {n} tnd {tnd}")
            np.testing.assert_array_equal(ora_sv, ref_sv, err_msg=f"sv case {n} tnd {tnd}")


@pytest.mark.parametrize("shape", ((3, 3, 1), (5, 5, 1), (7, 7, 2), (9, 31, 3), (40, 7, 5)))
def test_diffusers_and_smoothers_equal_reference_on_minimal_and_ragged_grids(shape):
    """The reference's HorizontalDiffusion / HorizontalSmoothing classes (2-D and 1-D variants,
    their own __call__: stencil + rim copies) against the oracle down to the smallest shape each
    scheme accepts (2 nb + 1 points along its axes)."""
    from oracle import dwarfs as od

    refload.install_framework()
    opts = refload.load("tasmania.framework.options")
    for name in ("second_order", "fourth_order"):
        refload.load("tasmania.dwarfs.subclasses.horizontal_diffusers." + name)
    for name in ("first_order", "second_order", "third_order"):
        refload.load("tasmania.dwarfs.subclasses.horizontal_smoothers." + name)
    hd = refload.load("tasmania.dwarfs.horizontal_diffusion")
    hsm = refload.load("tasmania.dwarfs.horizontal_smoothing")
    kw = dict(backend="numpy", backend_options=opts.BackendOptions(), storage_options=opts.StorageOptions())
    rng = np.random.default_rng(sum(shape))
    phi, base = rng.standard_normal(shape), rng.standard_normal(shape)
    depth = min(2, shape[2])
    checked = 0
    for sfx, axis in (("", None), ("_1dx", 0), ("_1dy", 1)):
        axes = (0, 1) if axis is None else (axis,)
        for order, name in ((2, "second_order"), (4, "fourth_order")):
            nb = order // 2
            if any(shape[a] < 2 * nb + 1 for a in axes):
                continue
            g = np.zeros(shape)
            g[...] = od.vertical_profile(0.5, 1.0, depth, shape[2])[None, None, :]
            origin = tuple(nb if a in axes else 0 for a in range(3))
            domain = tuple(n - 2 * nb if a in axes else n for a, n in enumerate(shape))
            obj = hd.HorizontalDiffusion.factory(name + sfx, shape, 0.7, 1.3, 0.5, 1.0, depth, **kw)
            np.testing.assert_array_equal(np.asarray(obj._gamma), g)
            for overwrite in (True, False):
                ref, ora = base.copy(), base.copy()
                obj(phi, ref, overwrite_output=overwrite)
                if axis is None:
                    od.diffusion(order, phi, g, ora, 0.7, 1.3, overwrite, origin, domain)
                else:
                    od.diffusion_1d(order, axis, phi, g, ora, (0.7, 1.3)[axis], overwrite, origin, domain)
                np.testing.assert_array_equal(ora, ref, err_msg=f"{name}{sfx} overwrite={overwrite}")
                checked += 1
        for order, name in ((1, "first_order"), (2, "second_order"), (3, "third_order")):
            if any(shape[a] < 2 * order + 1 for a in axes):
                continue
            g = np.zeros(shape)
            g[...] = od.vertical_profile(0.03, 0.24, depth, shape[2])[None, None, :]
            obj = hsm.HorizontalSmoothing.factory(name + sfx, shape, 0.03, 0.24, depth, **kw)
            ref, ora = base.copy(), base.copy()
            obj(phi, ref)
            if axis is None:
                od.horizontal_smoothing(order, phi, g, ora)
            else:
                od.horizontal_smoothing_1d(order, axis, phi, g, ora)
            np.testing.assert_array_equal(ora, ref, err_msg=f"{name}{sfx}")
            checked += 1
    assert checked >= 3


@pytest.mark.parametrize("dims", ((1, 1, 1), (2, 3, 1), (9, 4, 2), (5, 7, 64), (12, 10, 70)))
def test_k3_column_scans_equal_reference_from_one_level_to_deep_columns(dims):
    """IsentropicDiagnostics' numpy stencils (isentropic/dynamics/diagnostics.py:L319-L570:
    diagnostic_variables, montgomery, height, density_and_temperature) run in place against the
    oracle on a single column / single level and on columns deeper than the 64 levels the
    register-resident GPU scan covers -- bit for bit (both sides are numpy + glibc pow)."""
    refload.install_framework()
    diag = refload.load("tasmania.isentropic.dynamics.diagnostics")
    consts = {"pref": 1.0e5, "rd": 287.05, "g": 9.80665, "cp": 1004.0}
    D = diag.IsentropicDiagnostics
    mont = refload.numpy_stencil(D._montgomery_numpy, consts)
    dvar = refload.numpy_stencil(D._diagnostic_variables_numpy, consts)
    hgt = refload.numpy_stencil(D._height_numpy, consts)
    dat = refload.numpy_stencil(D._density_and_temperature_numpy, consts)
    nx, ny, nz = dims
    shape = (nx + 1, ny + 1, nz + 1)
    rng = np.random.default_rng(nx * 100 + nz)
    theta1d = np.linspace(400.0, 280.0, nz + 1)
    theta = np.zeros(shape)
    theta[:nx, :ny, :] = theta1d[None, None, :]
    hs = np.zeros(shape)
    hs[:nx, :ny, nz] = rng.uniform(0, 800, size=(nx, ny))
    s = rng.uniform(5, 60, size=shape)
    dz, pt = 120.0 / nz, 11868.9
    box = dict(origin=(0, 0, 0), domain=(nx, ny, nz + 1))

    ref, ora = np.zeros(shape), np.zeros(shape)
    mont(in_hs=hs, in_s=s, inout_mtg=ref, dz=dz, pt=pt, theta_s=theta1d[-1], **box)
    oi.montgomery(hs, s, ora, dz=dz, pt=pt, theta_s=theta1d[-1], **box)
    np.testing.assert_array_equal(ora, ref)

    r = [np.zeros(shape) for _ in range(4)]
    o = [np.zeros(shape) for _ in range(4)]
    dvar(in_theta=theta, in_hs=hs, in_s=s, inout_p=r[0], out_exn=r[1], inout_mtg=r[2], inout_h=r[3],
         dz=dz, pt=pt, **box)
    oi.diagnostic_variables(theta, hs, s, *o, dz=dz, pt=pt, **box)
    for a, b, n in zip(o, r, ("p", "exn", "mtg", "h")):
        np.testing.assert_array_equal(a, b, err_msg=n)
    np.testing.assert_array_equal(o[2], ora)  # the two Montgomery scans agree

    ref, ora = np.zeros(shape), np.zeros(shape)
    hgt(in_theta=theta, in_hs=hs, in_s=s, inout_h=ref, dz=dz, pt=pt, **box)
    oi.height(theta, hs, s, ora, dz=dz, pt=pt, **box)
    np.testing.assert_array_equal(ora, ref)

    rr, rt, orho, ot = (np.zeros(shape) for _ in range(4))
    dat(in_theta=theta, in_s=s, in_exn=r[1], in_h=r[3], out_rho=rr, out_t=rt, origin=(0, 0, 0),
        domain=(nx, ny, nz))
    oi.density_and_temperature(theta, s, o[1], o[3], orho, ot, origin=(0, 0, 0), domain=(nx, ny, nz))
    np.testing.assert_array_equal(orho, rr)
    np.testing.assert_array_equal(ot, rt)
