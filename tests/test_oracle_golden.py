# -*- coding: utf-8 -*-
"""Pin the oracle against fixtures produced by the reference's own numpy code
(tests/golden/generate_golden.py).  Bit-exact: same numpy, same operation order."""
import numpy as np
import pytest

from oracle import boundary as ob
from oracle import dwarfs as od
from oracle import isentropic as oi
from oracle.fluxes import EXTENT
from tests import helpers as hp

SCHEMES = ("upwind", "centered", "third_order_upwind", "fifth_order_upwind")


def eq(a, b):
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("moist", (False, True))
@pytest.mark.parametrize("tnd", (False, True))
def test_k1_k2(stencils_golden, scheme, moist, tnd):
    fx = stencils_golden
    nx, ny, nz = (int(v) for v in fx["dims"])
    dt, dx, dy, eps = fx["scalars"]
    e = EXTENT[scheme]
    origin, domain = (e, e, 0), (nx - 2 * e, ny - 2 * e, nz)
    shape = fx["s_now"].shape
    tag = f"{scheme}_m{int(moist)}_t{int(tnd)}"
    s_new = np.zeros(shape)
    sq_new = [np.zeros(shape) for _ in range(3)]
    kw = {}
    if moist:
        kw = dict(moist=True, sq_now=list(fx["sq_now"]), sq_int=list(fx["sq_int"]), sq_new=sq_new,
                  q_tnd=list(fx["q_tnd"]) if tnd else (None, None, None))
    oi.step_forward_euler(scheme, fx["s_now"], fx["s_int"], s_new, fx["u_int"], fx["v_int"],
                          dt=dt, dx=dx, dy=dy, origin=origin, domain=domain,
                          s_tnd=fx["s_tnd"] if tnd else None, **kw)
    eq(s_new, fx[f"k1_{tag}_s_new"])
    if moist:
        eq(np.stack(sq_new), fx[f"k1_{tag}_sq_new"])
        return
    su_new, sv_new = np.zeros(shape), np.zeros(shape)
    oi.step_forward_euler_momentum(
        scheme, fx["s_now"], fx["s_new_in"], fx["u_int"], fx["v_int"], fx["su_now"], fx["su_int"],
        su_new, fx["sv_now"], fx["sv_int"], sv_new, fx["mtg_now"], fx["mtg_new"], dt=dt, dx=dx,
        dy=dy, eps=eps, origin=origin, domain=domain,
        su_tnd=fx["su_tnd"] if tnd else None, sv_tnd=fx["sv_tnd"] if tnd else None)
    eq(su_new, fx[f"k2_{tag}_su_new"])
    eq(sv_new, fx[f"k2_{tag}_sv_new"])


def test_k3_diagnostics(stencils_golden):
    fx = stencils_golden
    nx, ny, nz = (int(v) for v in fx["dims"])
    dz, pt, theta_s = fx["k3_scalars"]
    shape = fx["k3_s"].shape
    box = dict(origin=(0, 0, 0), domain=(nx, ny, nz + 1))
    mtg = np.zeros(shape)
    oi.montgomery(fx["k3_hs"], fx["k3_s"], mtg, dz=dz, pt=pt, theta_s=theta_s, **box)
    eq(mtg, fx["k3_mtg"])
    p, exn, mtg2, h = (np.zeros(shape) for _ in range(4))
    oi.diagnostic_variables(fx["k3_theta"], fx["k3_hs"], fx["k3_s"], p, exn, mtg2, h, dz=dz, pt=pt,
                            **box)
    for a, n in ((p, "p"), (exn, "exn"), (mtg2, "mtg2"), (h, "h")):
        eq(a, fx["k3_" + n])
    h2 = np.zeros(shape)
    oi.height(fx["k3_theta"], fx["k3_hs"], fx["k3_s"], h2, dz=dz, pt=pt, **box)
    eq(h2, fx["k3_h2"])
    rho, t = np.zeros(shape), np.zeros(shape)
    oi.density_and_temperature(fx["k3_theta"], fx["k3_s"], exn, h, rho, t, origin=(0, 0, 0),
                               domain=(nx, ny, nz))
    eq(rho, fx["k3_rho"])
    eq(t, fx["k3_t"])


def test_k4_k7(stencils_golden):
    fx = stencils_golden
    nx, ny, nz = (int(v) for v in fx["dims"])
    shape = fx["s_now"].shape
    u, v = np.zeros(shape), np.zeros(shape)
    od.get_velocity_components(nx, ny, nz, fx["s_now"], fx["su_now"], fx["sv_now"], u, v)
    eq(u, fx["k4_u"])
    eq(v, fx["k4_v"])
    du, dv = np.zeros(shape), np.zeros(shape)
    od.momenta(fx["s_now"], fx["u_int"], fx["v_int"], du, dv, (0, 0, 0), (nx, ny, nz))
    eq(du, fx["k4_du"])
    eq(dv, fx["k4_dv"])
    sq, q = np.zeros(shape), np.zeros(shape)
    od.density(fx["s_now"], fx["k7_q"], sq, (0, 0, 0), (nx, ny, nz))
    eq(sq, fx["k7_sq"])
    od.mass_fraction(fx["s_now"], fx["k7_sq_in"], q, (0, 0, 0), (nx, ny, nz))
    eq(q, fx["k7_q_out"])


def test_k5_relaxed(stencils_golden):
    fx = stencils_golden
    nx, ny, nz = (int(v) for v in fx["dims"])
    gamma = ob.relaxed_gamma(nx, ny, nz, 3, 6)
    eq(gamma, fx["k5_gamma"])
    phi = fx["k5_phi"].copy()
    ob.irelax(gamma, fx["k5_phi_ref"], phi, (0, 0, 0), (nx, ny, nz))
    eq(phi, fx["k5_irelax"])
    out = np.zeros_like(phi)
    ob.relax(gamma, fx["k5_phi"], fx["k5_phi_ref"], out, (0, 0, 0), (nx + 1, ny, nz))
    eq(out, fx["k5_relax"])


def test_k6_rayleigh(stencils_golden):
    fx = stencils_golden
    shape = fx["k5_phi"].shape
    depth, cmax, dt = fx["k6_params"]
    r = od.rayleigh_coefficient(fx["k6_z"], fx["k6_zhl"][0], int(depth), cmax, shape[2])
    rmat = np.zeros(shape)
    rmat[...] = r[None, None, :]
    eq(rmat, fx["k6_rmat"])
    out = np.zeros(shape)
    od.damping(fx["k5_phi"], fx["k5_phi_ref"], fx["s_now"], rmat, out, dt, (0, 0, 0), shape)
    eq(out, fx["k6_out"])


@pytest.mark.parametrize("order", (2, 4))
def test_k8_diffusion(stencils_golden, order):
    fx = stencils_golden
    shape = fx["k5_phi"].shape
    _, dx, dy, _ = fx["scalars"]
    nb = order // 2
    gamma = np.zeros(shape)
    gamma[...] = od.vertical_profile(0.5, 1.0, 3, shape[2])[None, None, :]
    eq(gamma, fx[f"k8_{order}_gamma"])
    origin = (nb, nb, 0)
    domain = (shape[0] - 2 * nb, shape[1] - 2 * nb, shape[2])
    tnd = np.zeros(shape)
    od.diffusion(order, fx["k5_phi"], gamma, tnd, dx, dy, True, origin, domain)
    eq(tnd, fx[f"k8_{order}_tnd"])
    acc = fx["k5_phi_ref"].copy()
    od.diffusion(order, fx["k5_phi"], gamma, acc, dx, dy, False, origin, domain)
    eq(acc, fx[f"k8_{order}_acc"])


@pytest.mark.parametrize("order", (1, 2, 3))
def test_k9_smoothing(stencils_golden, order):
    fx = stencils_golden
    shape = fx["k5_phi"].shape
    gamma = np.zeros(shape)
    gamma[...] = od.vertical_profile(0.03, 0.24, 3, shape[2])[None, None, :]
    eq(gamma, fx[f"k9_{order}_gamma"])
    out = np.zeros(shape)
    od.horizontal_smoothing(order, fx["k5_phi"], gamma, out)
    eq(out, fx[f"k9_{order}_out"])


@pytest.mark.parametrize("ax", ("x", "y"))
@pytest.mark.parametrize("tag", ("row", "grid"))
def test_k8_k9_one_dimensional_variants(stencils_1d_golden, ax, tag):
    """The reference's ..._1dx / ..._1dy diffusers and smoothers (run in place by
    golden/generate_golden.py stencils_1d) against the oracle, bit for bit."""
    fx = stencils_1d_golden
    axis = 0 if ax == "x" else 1
    phi = fx[f"{ax}_{tag}_phi"]
    shape = phi.shape
    h = fx["scalars"][axis]
    gamma = np.zeros(shape)
    gamma[...] = od.vertical_profile(0.5, 1.0, 3, shape[2])[None, None, :]
    for order in (2, 4):
        nb = order // 2
        origin = (nb, 0, 0) if axis == 0 else (0, nb, 0)
        domain = tuple(n - 2 * nb if a == axis else n for a, n in enumerate(shape))
        tnd = np.zeros(shape)
        od.diffusion_1d(order, axis, phi, gamma, tnd, h, True, origin, domain)
        eq(tnd, fx[f"k8_{order}_{ax}_{tag}_tnd"])
        acc = fx[f"{ax}_{tag}_base"].copy()
        od.diffusion_1d(order, axis, phi, gamma, acc, h, False, origin, domain)
        eq(acc, fx[f"k8_{order}_{ax}_{tag}_acc"])
    gamma[...] = od.vertical_profile(0.03, 0.24, 3, shape[2])[None, None, :]
    for order in (1, 2, 3):
        out = np.zeros(shape)
        od.horizontal_smoothing_1d(order, axis, phi, gamma, out)
        eq(out, fx[f"k9_{order}_{ax}_{tag}_out"])


def test_thomas_global_stencil(stencils_1d_golden):
    """stencil_definitions/cla.py:thomas_numpy run in place (zero-pivot columns included)."""
    from oracle import isentropic_physics as op

    fx = stencils_1d_golden
    box = [int(v) for v in fx["thomas_box"]]
    x = np.zeros(fx["thomas_a"].shape)
    op.thomas(fx["thomas_a"], fx["thomas_b"], fx["thomas_c"], fx["thomas_d"], x, box[:3], box[3:])
    eq(x, fx["thomas_x"])
    # the specialised solver of the implicit vertical advection (b == 1) agrees with it
    a, c, d = fx["thomas_a"][1:, :, 1:8], fx["thomas_c"][1:, :, 1:8], fx["thomas_d"][1:, :, 1:8]
    y = np.zeros(fx["thomas_a"].shape)
    op.thomas(fx["thomas_a"], np.ones_like(fx["thomas_b"]), fx["thomas_c"], fx["thomas_d"], y,
              box[:3], box[3:])
    eq(op._thomas(a, c, d), y[1:, :, 1:8])


def test_hyperdiffusion_global_stencil(stencils_1d_golden):
    """stencil_definitions/diffusion.py:diffusion_numpy run in place."""
    fx = stencils_1d_golden
    box = [int(v) for v in fx["hyper_box"]]
    out = np.zeros(fx["hyper_phi"].shape)
    od.hyperdiffusion(fx["hyper_phi"], out, float(fx["hyper_alpha"]), box[:3], box[3:])
    eq(out, fx["hyper_out"])


def test_k12_elementwise(stencils_golden):
    fx = stencils_golden
    nx, ny, nz = (int(v) for v in fx["dims"])
    a, b, c = fx["k12_a"], fx["k12_b"], fx["k12_c"]
    box = (slice(1, nx - 1), slice(2, ny - 1), slice(0, nz))
    E = od.ELEMENTWISE

    def chk(name, val, base=None):
        exp = np.zeros_like(a) if base is None else base.copy()
        exp[box] = val
        eq(exp, fx["k12_" + name])

    chk("abs", E["abs"](a[box]))
    chk("add", E["add"](a[box], b[box]))
    chk("addsub", E["addsub"](a[box], b[box], c[box]))
    chk("clip", E["clip"](a[box]))
    chk("fma", E["fma"](a[box], b[box], 0.37))
    chk("mul", E["mul"](a[box], b[box]))
    chk("scale", E["scale"](a[box], -1.7))
    chk("sub", E["sub"](a[box], b[box]))
    chk("copy", a[box])
    chk("copychange", -a[box])
    chk("sts_rk2_0", E["sts_rk2_0"](a[box], b[box], c[box], 0.8))
    chk("sts_rk3ws_0", E["sts_rk3ws_0"](a[box], b[box], c[box], 0.8))
    chk("iabs", np.abs(a[box]), a)
    chk("iadd", a[box] + b[box], a)
    chk("iaddsub", a[box] + (b[box] - c[box]), a)
    chk("iclip", np.where(a[box] > 0, a[box], 0), a)
    chk("imul", a[box] * b[box], a)
    chk("iscale", a[box] * 2.5, a)
    chk("isub", a[box] - b[box], a)


@pytest.mark.parametrize(
    "case", ("isen_dry_rk3_5th", "isen_dry_rk3_3rd", "isen_dry_rk3_cen", "isen_dry_fe_upw",
             "isen_dry_rk3_5th_periodic", "isen_dry_rk3_3rd_periodic", "isen_dry_fe_cen_periodic")
)
def test_dry_dycore_orchestration(case):
    """The oracle's restated orchestration vs. the reference's own classes driving the
    reference's own ``stage_array_call_dry`` over several steps (relaxed boundaries, and the
    reference's Periodic class on its numerical grid)."""
    fx = hp.load(case)
    final, stage0, _ = hp.oracle_dry_run(fx)
    nx, ny, nz = (int(v) for v in fx["dims"][:3])
    for n in (hp.S, hp.SU, hp.SV, hp.U, hp.V):
        eq(stage0[n][: nx + 1, : ny + 1, :nz], fx["stage0_" + n][: nx + 1, : ny + 1, :nz])
    for n in (hp.S, hp.SU, hp.SV, hp.U, hp.V, hp.MTG, hp.P, hp.EXN, hp.H):
        eq(final[n], fx["final_" + n])


@pytest.mark.parametrize("case", ("isen_moist_rk3_5th", "isen_moist_fe_3rd", "isen_moist_rk3_5th_periodic"))
def test_moist_dycore_orchestration(case):
    """Moist stage (tracer densities -> K1 with tracers -> mass fractions -> boundary -> damping
    -> velocities) against the reference's own ``stage_array_call_moist``."""
    fx = hp.load(case)
    final, stage0, _ = hp.oracle_dry_run(fx)
    nx, ny, nz = (int(v) for v in fx["dims"][:3])
    qn = (oi.MFWV, oi.MFCW, oi.MFPW)
    for n in (hp.S, hp.SU, hp.SV, hp.U, hp.V) + qn:
        eq(stage0[n][: nx + 1, : ny + 1, :nz], fx["stage0_" + n][: nx + 1, : ny + 1, :nz])
    for n in (hp.S, hp.SU, hp.SV, hp.U, hp.V, hp.MTG, hp.P, hp.EXN, hp.H) + qn:
        eq(final[n], fx["final_" + n])
    assert np.abs(final[oi.MFCW]).max() > 0 and np.abs(final[oi.MFPW]).max() > 0


# ------------------------------------------------------------------ K11 Kessler
def test_kessler_family_bitwise():
    """Every K11 stencil of the oracle against the reference's own numpy definitions
    (tests/golden/kessler.npz), bit for bit."""
    from oracle import microphysics as om
    from tests.kessler_cases import cases

    fx = hp.load("kessler")
    n = 0
    for tag, name, got, want in cases(fx, om, lambda a: np.array(a, copy=True), np.asarray, np.zeros):
        np.testing.assert_array_equal(got, want, err_msg=f"{tag}:{name}")
        n += 1
    assert n >= 60


def _vertical_advection_cases(fx):
    """(key prefix, scheme, staggered, moist, overwrite) of tests/golden/isentropic_physics.npz"""
    from oracle import isentropic_physics as va

    for scheme in va.EXTENT:
        for z in (0, 1):
            for m in (0, 1):
                for ow in (1, 0):
                    yield f"{scheme}_z{z}_m{m}_o{ow}_", scheme, bool(z), bool(m), bool(ow)


def test_vertical_advection_bitwise():
    """the oracle's vertical advection against the reference's own numpy stencil
    (vertical_advection.py:L271-L386) for the four flux schemes, velocity on main / interface
    levels, dry / moist, overwrite on / off -- bit for bit, on the WHOLE output storages."""
    from oracle import isentropic_physics as va

    fx = hp.load("isentropic_physics")
    nx, ny, nz = (int(v) for v in fx["dims"])
    dz = float(fx["dz"][0])
    n = 0
    for prefix, scheme, stg, moist, ow in _vertical_advection_cases(fx):
        names = ("s", "su", "sv") + (("qv", "qc", "qr") if moist else ())
        outs = {k: (np.full(fx["in_s"].shape, 7.0) if ow else fx["prev_" + k].copy()) for k in names}
        kw = dict(dz=dz, origin=(0, 0, 0), domain=(nx, ny, nz))
        for k in names:
            kw["ow_out_" + k] = ow
        if moist:
            for k in ("qv", "qc", "qr"):
                kw["in_" + k] = fx["in_" + k]
                kw["out_" + k] = outs[k]
        va.vertical_advection(scheme, stg, fx["in_w"], fx["in_s"], fx["in_su"], fx["in_sv"],
                              outs["s"], outs["su"], outs["sv"], **kw)
        for k in names:
            np.testing.assert_array_equal(outs[k], fx[prefix + k], err_msg=prefix + k)
            n += 1
    assert n == 4 * 2 * (3 + 6) * 2


def test_coriolis_bitwise():
    """coriolis.py:L166-L186 on the interior box, written and accumulated."""
    from oracle import isentropic_physics as va

    fx = hp.load("isentropic_physics")
    nx, ny, nz = (int(v) for v in fx["dims"])
    for owu, owv in ((True, True), (False, True), (False, False)):
        tu, tv = fx["prev_su"].copy(), fx["prev_sv"].copy()
        va.coriolis(fx["in_su"], fx["in_sv"], tu, tv, f=float(fx["f"][0]), ow_tnd_su=owu, ow_tnd_sv=owv,
                    origin=(2, 2, 0), domain=(nx - 4, ny - 4, nz))
        np.testing.assert_array_equal(tu, fx[f"coriolis_o{int(owu)}{int(owv)}_su"])
        np.testing.assert_array_equal(tv, fx[f"coriolis_o{int(owu)}{int(owv)}_sv"])


def test_smagorinsky_bitwise():
    """Smagorinsky2d and IsentropicSmagorinsky against the reference's own numpy
    stencils (physics/turbulence.py:L165-L229, isentropic/physics/turbulence.py:L99-L125)."""
    from oracle import isentropic_physics as va

    fx = hp.load("isentropic_physics")
    nx, ny, nz = (int(v) for v in fx["dims"])
    dx, dy, cs = (float(v) for v in fx["smag"])
    for nb, ow in ((2, True), (3, False)):
        box = dict(origin=(nb, nb, 0), domain=(nx - 2 * nb, ny - 2 * nb, nz))
        a, b = fx["prev_su"].copy(), fx["prev_sv"].copy()
        va.smagorinsky(fx["in_u"], fx["in_v"], a, b, dx=dx, dy=dy, cs=cs, ow_out_u_tnd=ow,
                       ow_out_v_tnd=not ow, **box)
        np.testing.assert_array_equal(a, fx[f"smag2d_nb{nb}_u"])
        np.testing.assert_array_equal(b, fx[f"smag2d_nb{nb}_v"])
        a, b = fx["prev_su"].copy(), fx["prev_sv"].copy()
        va.smagorinsky(fx["in_su"], fx["in_sv"], a, b, in_s=fx["in_s"], dx=dx, dy=dy, cs=cs,
                       ow_out_u_tnd=ow, ow_out_v_tnd=not ow, **box)
        np.testing.assert_array_equal(a, fx[f"smagisen_nb{nb}_su"])
        np.testing.assert_array_equal(b, fx[f"smagisen_nb{nb}_sv"])


def test_implicit_vertical_advection_bitwise():
    """the Crank-Nicolson vertical advection (tridiagonal set-up + Thomas algorithm,
    implicit_vertical_advection.py:L221-L336, cla.py:L42-L108) against the reference's own numpy
    code: velocity on main / interface levels, dry / moist."""
    from oracle import isentropic_physics as va

    fx = hp.load("isentropic_physics")
    nx, ny, nz = (int(v) for v in fx["dims"])
    for z in (0, 1):
        for m in (0, 1):
            names = ("s", "su", "sv") + (("qv", "qc", "qr") if m else ())
            outs = {n: fx["prev_" + n].copy() for n in names}
            kw = dict(gamma=float(fx["gamma"][0]), origin=(0, 0, 0), domain=(nx, ny, nz))
            if m:
                for n in ("qv", "qc", "qr"):
                    kw["in_" + n], kw["out_" + n] = fx["in_" + n], outs[n]
            va.implicit_vertical_advection(bool(z), fx["in_w_implicit"], fx["in_s"], fx["in_su"],
                                           fx["in_sv"], outs["s"], outs["su"], outs["sv"], **kw)
            for n in names:
                np.testing.assert_array_equal(outs[n], fx[f"implicit_z{z}_m{m}_{n}"], err_msg=f"{z}{m}{n}")


def test_implicit_vertical_advection_tendency_bitwise():
    """The prognostic variant (implicit_vertical_advection.py:L793-L919)."""
    from oracle import isentropic_physics as va

    fx = hp.load("isentropic_physics")
    nx, ny, nz = (int(v) for v in fx["dims"])
    for z, m in ((0, 1), (1, 0)):
        names = ("s", "su", "sv") + (("qv", "qc", "qr") if m else ())
        outs = {n: fx["prev_" + n].copy() for n in names}
        kw = dict(dt=7.5, gamma=float(fx["gamma"][0]), origin=(0, 0, 0), domain=(nx, ny, nz))
        if m:
            for n in ("qv", "qc", "qr"):
                kw["in_" + n], kw["tnd_" + n] = fx["in_" + n], outs[n]
        va.implicit_vertical_advection_tendency(bool(z), fx["in_w_implicit"], fx["in_s"], fx["in_su"],
                                                fx["in_sv"], outs["s"], outs["su"], outs["sv"], **kw)
        for n in names:
            np.testing.assert_array_equal(outs[n], fx[f"implicit_tnd_z{z}_m{m}_{n}"], err_msg=f"{z}{m}{n}")
