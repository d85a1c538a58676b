# -*- coding: utf-8 -*-
"""Parity of every stand-alone CUDA stencil (through the registry -> ctypes -> C ABI path)
against (a) the golden fixtures produced by the reference's own numpy code and (b) the oracle
on fresh seeded inputs with ragged sizes.

Tolerances: kernels without libm calls are compiled with -fmad=false / IEEE division and must
be BIT-EXACT (rtol 0).  Kernels calling ``pow`` (K3) are held to 1e-13 relative (CUDA libm vs
glibc may differ in the last ulp) -- ten times tighter than north_star's 1e-12.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import tasmania_b200 as tb  # noqa: E402
from oracle import boundary as ob  # noqa: E402
from oracle import dwarfs as od  # noqa: E402
from oracle import isentropic as oi  # noqa: E402
from oracle.fluxes import EXTENT  # noqa: E402
from tasmania_b200.framework import BackendOptions  # noqa: E402
from tasmania_b200.stencils import ADVECTION, FLUX  # noqa: E402
from tests import helpers as hp  # noqa: E402

SCHEMES = ("upwind", "centered", "third_order_upwind", "fifth_order_upwind")
POW_RTOL = 1e-13


def dev(a):
    return tb.as_storage(np.asarray(a))


def stencil(name, **externals):
    return tb.compile_stencil(name, backend_options=BackendOptions(externals=externals))


def eq(a, b):
    np.testing.assert_array_equal(tb.to_numpy(a), b)


def close(a, b, rtol=POW_RTOL):
    a = tb.to_numpy(a)
    assert hp.relerr(a, b) <= rtol, hp.relerr(a, b)


@pytest.fixture(scope="module")
def fx():
    return hp.load("stencils")


@pytest.mark.parametrize("scheme", SCHEMES)
@pytest.mark.parametrize("moist", (False, True))
@pytest.mark.parametrize("tnd", (False, True))
def test_k1_k2_golden(fx, scheme, moist, tnd):
    nx, ny, nz = (int(v) for v in fx["dims"])
    dt, dx, dy, eps = (float(v) for v in fx["scalars"])
    e = EXTENT[scheme]
    origin, domain = (e, e, 0), (nx - 2 * e, ny - 2 * e, nz)
    shape = fx["s_now"].shape
    tag = f"{scheme}_m{int(moist)}_t{int(tnd)}"
    ext = dict(flux_dry=FLUX[scheme], flux_moist=FLUX[scheme], extent=e, moist=moist)
    k1 = stencil("step_forward_euler", **ext)
    s_new = tb.zeros(shape)
    sq_new = [tb.zeros(shape) for _ in range(3)]
    kw = {}
    if moist:
        for n, t in zip(("sqv", "sqc", "sqr"), range(3)):
            kw[n + "_now"], kw[n + "_int"], kw[n + "_new"] = dev(fx["sq_now"][t]), dev(fx["sq_int"][t]), sq_new[t]
        if tnd:
            kw.update(qv_tnd=dev(fx["q_tnd"][0]), qc_tnd=dev(fx["q_tnd"][1]), qr_tnd=dev(fx["q_tnd"][2]))
    k1(s_now=dev(fx["s_now"]), s_int=dev(fx["s_int"]), s_new=s_new, u_int=dev(fx["u_int"]),
       v_int=dev(fx["v_int"]), su_int=dev(fx["su_int"]), sv_int=dev(fx["sv_int"]),
       s_tnd=dev(fx["s_tnd"]) if tnd else None, dt=dt, dx=dx, dy=dy, origin=origin, domain=domain,
       exec_info=None, validate_args=False, **kw)
    eq(s_new, fx[f"k1_{tag}_s_new"])
    if moist:
        for t in range(3):
            eq(sq_new[t], fx[f"k1_{tag}_sq_new"][t])
        return
    k2 = stencil("step_forward_euler_momentum", **ext)
    su_new, sv_new = tb.zeros(shape), tb.zeros(shape)
    k2(s_now=dev(fx["s_now"]), s_int=dev(fx["s_int"]), s_new=dev(fx["s_new_in"]),
       u_int=dev(fx["u_int"]), v_int=dev(fx["v_int"]), su_now=dev(fx["su_now"]),
       su_int=dev(fx["su_int"]), su_new=su_new, sv_now=dev(fx["sv_now"]), sv_int=dev(fx["sv_int"]),
       sv_new=sv_new, mtg_now=dev(fx["mtg_now"]), mtg_new=dev(fx["mtg_new"]),
       su_tnd=dev(fx["su_tnd"]) if tnd else None, sv_tnd=dev(fx["sv_tnd"]) if tnd else None,
       dt=dt, dx=dx, dy=dy, eps=eps, origin=origin, domain=domain)
    eq(su_new, fx[f"k2_{tag}_su_new"])
    eq(sv_new, fx[f"k2_{tag}_sv_new"])


def test_k3_golden(fx):
    nx, ny, nz = (int(v) for v in fx["dims"])
    dz, pt, theta_s = (float(v) for v in fx["k3_scalars"])
    shape = fx["k3_s"].shape
    box = dict(origin=(0, 0, 0), domain=(nx, ny, nz + 1))
    c = oi.CONSTANTS
    mtg = tb.zeros(shape)
    stencil("montgomery", **c)(in_hs=dev(fx["k3_hs"]), in_s=dev(fx["k3_s"]), inout_mtg=mtg, dz=dz,
                               pt=pt, theta_s=theta_s, **box)
    close(mtg, fx["k3_mtg"])
    p, exn, mtg2, h = (tb.zeros(shape) for _ in range(4))
    stencil("diagnostic_variables", **c)(
        in_theta=dev(fx["k3_theta"]), in_hs=dev(fx["k3_hs"]), in_s=dev(fx["k3_s"]), inout_p=p,
        out_exn=exn, inout_mtg=mtg2, inout_h=h, dz=dz, pt=pt, **box)
    eq(p, fx["k3_p"])  # the pressure scan has no libm call: bit-exact
    for a, n in ((exn, "exn"), (mtg2, "mtg2"), (h, "h")):
        close(a, fx["k3_" + n])
    h2 = tb.zeros(shape)
    stencil("height", **c)(in_theta=dev(fx["k3_theta"]), in_hs=dev(fx["k3_hs"]),
                           in_s=dev(fx["k3_s"]), inout_h=h2, dz=dz, pt=pt, **box)
    close(h2, fx["k3_h2"])
    rho, t = tb.zeros(shape), tb.zeros(shape)
    stencil("density_and_temperature", **c)(
        in_theta=dev(fx["k3_theta"]), in_s=dev(fx["k3_s"]), in_exn=dev(fx["k3_exn"]),
        in_h=dev(fx["k3_h"]), out_rho=rho, out_t=t, origin=(0, 0, 0), domain=(nx, ny, nz))
    eq(rho, fx["k3_rho"])
    eq(t, fx["k3_t"])


def test_k4_k7_golden(fx):
    nx, ny, nz = (int(v) for v in fx["dims"])
    shape = fx["s_now"].shape
    s, su, sv = dev(fx["s_now"]), dev(fx["su_now"]), dev(fx["sv_now"])
    u, v = tb.zeros(shape), tb.zeros(shape)
    stencil("velocity_x", staggering=True)(in_d=s, in_du=su, out_u=u, origin=(1, 0, 0),
                                           domain=(nx - 1, ny, nz))
    stencil("velocity_y", staggering=True)(in_d=s, in_dv=sv, out_v=v, origin=(0, 1, 0),
                                           domain=(nx, ny - 1, nz))
    eq(u, fx["k4_u"])
    eq(v, fx["k4_v"])
    du, dv = tb.zeros(shape), tb.zeros(shape)
    stencil("momenta", staggering=True)(in_d=s, in_u=dev(fx["u_int"]), in_v=dev(fx["v_int"]),
                                        out_du=du, out_dv=dv, origin=(0, 0, 0), domain=(nx, ny, nz))
    eq(du, fx["k4_du"])
    eq(dv, fx["k4_dv"])
    sq, q = tb.zeros(shape), tb.zeros(shape)
    stencil("density", clipping=True)(in_d=s, in_q=dev(fx["k7_q"]), out_dq=sq, origin=(0, 0, 0),
                                      domain=(nx, ny, nz))
    eq(sq, fx["k7_sq"])
    stencil("mass_fraction", clipping=True)(in_d=s, in_dq=dev(fx["k7_sq_in"]), out_q=q,
                                            origin=(0, 0, 0), domain=(nx, ny, nz))
    eq(q, fx["k7_q_out"])


def test_k5_k6_golden(fx):
    nx, ny, nz = (int(v) for v in fx["dims"])
    shape = fx["k5_phi"].shape
    from tasmania_b200.boundary import Relaxed

    hb = Relaxed(nx, ny, nz, 3, nr=6)
    eq(hb._gamma, fx["k5_gamma"])  # index masks: bit-exact
    phi = dev(fx["k5_phi"])
    stencil("irelax")(in_gamma=dev(fx["k5_gamma"]), in_phi_ref=dev(fx["k5_phi_ref"]), inout_phi=phi,
                      origin=(0, 0, 0), domain=(nx, ny, nz))
    eq(phi, fx["k5_irelax"])
    out = tb.zeros(shape)
    stencil("relax")(in_gamma=hb._gamma, in_phi=dev(fx["k5_phi"]), in_phi_ref=dev(fx["k5_phi_ref"]),
                     out_phi=out, origin=(0, 0, 0), domain=(nx + 1, ny, nz))
    eq(out, fx["k5_relax"])
    depth, cmax, dt = fx["k6_params"]
    out = tb.zeros(shape)
    stencil("damping")(in_phi_now=dev(fx["k5_phi"]), in_phi_new=dev(fx["k5_phi_ref"]),
                       in_phi_ref=dev(fx["s_now"]), in_rmat=dev(fx["k6_rmat"]), out_phi=out,
                       dt=float(dt), origin=(0, 0, 0), domain=shape)
    eq(out, fx["k6_out"])


@pytest.mark.parametrize("order,name", ((2, "second_order"), (4, "fourth_order")))
def test_k8_golden(fx, order, name):
    from tasmania_b200.dwarfs import HorizontalDiffusion

    shape = fx["k5_phi"].shape
    _, dx, dy, _ = (float(v) for v in fx["scalars"])
    obj = HorizontalDiffusion.factory(name, shape, dx, dy, 0.5, 1.0, 3)
    eq(obj._gamma, fx[f"k8_{order}_gamma"])
    tnd = tb.zeros(shape)
    obj(dev(fx["k5_phi"]), tnd, overwrite_output=True)
    eq(tnd, fx[f"k8_{order}_tnd"])
    acc = dev(fx["k5_phi_ref"])
    obj(dev(fx["k5_phi"]), acc, overwrite_output=False)
    eq(acc, fx[f"k8_{order}_acc"])


@pytest.mark.parametrize("order,name", ((1, "first_order"), (2, "second_order"), (3, "third_order")))
def test_k9_golden(fx, order, name):
    from tasmania_b200.dwarfs import HorizontalSmoothing

    shape = fx["k5_phi"].shape
    obj = HorizontalSmoothing.factory(name, shape, 0.03, 0.24, 3)
    eq(obj._gamma, fx[f"k9_{order}_gamma"])
    out = tb.zeros(shape)
    obj(dev(fx["k5_phi"]), out)
    eq(out, fx[f"k9_{order}_out"])


def test_k12_golden(fx):
    nx, ny, nz = (int(v) for v in fx["dims"])
    a, b, c = fx["k12_a"], fx["k12_b"], fx["k12_c"]
    shape = a.shape
    box = dict(origin=(1, 2, 0), domain=(nx - 2, ny - 3, nz))

    def out_of(name, **kw):
        o = tb.zeros(shape)
        kw = {k: (dev(v) if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
        out_key = {"copy": "dst", "copychange": "dst", "abs": "out_field", "clip": "out_field",
                   "scale": "out_a", "addsub": "out_d", "sts_rk2_0": "out_field",
                   "sts_rk3ws_0": "out_field"}.get(name, "out_c")
        stencil(name)(**kw, **{out_key: o}, **box)
        eq(o, fx["k12_" + name])

    out_of("abs", in_field=a)
    out_of("add", in_a=a, in_b=b)
    out_of("addsub", in_a=a, in_b=b, in_c=c)
    out_of("clip", in_field=a)
    out_of("fma", in_a=a, in_b=b, f=0.37)
    out_of("mul", in_a=a, in_b=b)
    out_of("scale", in_a=a, f=-1.7)
    out_of("sub", in_a=a, in_b=b)
    out_of("copy", src=a)
    out_of("copychange", src=a)
    out_of("sts_rk2_0", in_field=a, in_field_prv=b, in_tnd=c, dt=0.8)
    out_of("sts_rk3ws_0", in_field=a, in_field_prv=b, in_tnd=c, dt=0.8)
    for name, kw in (("iabs", {}), ("iadd", dict(in_b=b)), ("iaddsub", dict(in_b=b, in_c=c)),
                     ("iclip", {}), ("imul", dict(in_b=b)), ("iscale", dict(f=2.5)),
                     ("isub", dict(in_b=b))):
        io = dev(a)
        key = "inout_field" if name in ("iabs", "iclip") else "inout_a"
        kw = {k: (dev(v) if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
        stencil(name)(**{key: io}, **kw, **box)
        eq(io, fx["k12_" + name])


# ------------------------------------------------------------------ oracle, ragged sizes
@pytest.mark.parametrize("shape", ((7, 9, 1), (33, 17, 3), (70, 41, 5), (130, 67, 2)))
@pytest.mark.parametrize("scheme", SCHEMES)
def test_k1_k2_oracle_ragged(shape, scheme):
    rng = np.random.default_rng(hash((shape, scheme)) % (2**32))
    nx, ny, nz = shape[0] - 1, shape[1] - 1, shape[2]
    e = EXTENT[scheme]
    f = lambda lo, hi: rng.uniform(lo, hi, size=shape)  # noqa: E731
    s_now, s_int, u, v = f(10, 1000), f(10, 1000), f(-50, 50), f(-50, 50)
    su_now, su_int, sv_now, sv_int = f(-5e3, 5e3), f(-5e3, 5e3), f(-5e3, 5e3), f(-5e3, 5e3)
    mtg_now, mtg_new, s_new_in = f(2.9e5, 3.1e5), f(2.9e5, 3.1e5), f(10, 1000)
    dt, dx, dy, eps = 2.5, 1234.5, 987.6, 0.5
    origin, domain = (e, e, 0), (nx - 2 * e, ny - 2 * e, nz)
    exp_s = np.zeros(shape)
    oi.step_forward_euler(scheme, s_now, s_int, exp_s, u, v, dt=dt, dx=dx, dy=dy, origin=origin,
                          domain=domain)
    exp_su, exp_sv = np.zeros(shape), np.zeros(shape)
    oi.step_forward_euler_momentum(scheme, s_now, s_new_in, u, v, su_now, su_int, exp_su, sv_now,
                                   sv_int, exp_sv, mtg_now, mtg_new, dt=dt, dx=dx, dy=dy, eps=eps,
                                   origin=origin, domain=domain)
    ext = dict(flux_dry=FLUX[scheme], flux_moist=FLUX[scheme], extent=e, moist=False)
    s_new = tb.zeros(shape)
    stencil("step_forward_euler", **ext)(
        s_now=dev(s_now), s_int=dev(s_int), s_new=s_new, u_int=dev(u), v_int=dev(v), dt=dt, dx=dx,
        dy=dy, origin=origin, domain=domain)
    eq(s_new, exp_s)
    su_new, sv_new = tb.zeros(shape), tb.zeros(shape)
    stencil("step_forward_euler_momentum", **ext)(
        s_now=dev(s_now), s_new=dev(s_new_in), u_int=dev(u), v_int=dev(v), su_now=dev(su_now),
        su_int=dev(su_int), su_new=su_new, sv_now=dev(sv_now), sv_int=dev(sv_int), sv_new=sv_new,
        mtg_now=dev(mtg_now), mtg_new=dev(mtg_new), dt=dt, dx=dx, dy=dy, eps=eps, origin=origin,
        domain=domain)
    eq(su_new, exp_su)
    eq(sv_new, exp_sv)


@pytest.mark.parametrize("shape", ((9, 9, 1), (66, 21, 4), (131, 75, 3)))
def test_diffusion_smoothing_oracle_ragged(shape):
    from tasmania_b200.dwarfs import HorizontalDiffusion, HorizontalSmoothing

    rng = np.random.default_rng(7 + shape[0])
    phi = rng.standard_normal(shape)
    for order, name in ((2, "second_order"), (4, "fourth_order")):
        nb = order // 2
        g = np.zeros(shape)
        g[...] = od.vertical_profile(0.5, 1.0, min(2, shape[2]), shape[2])[None, None, :]
        exp = np.zeros(shape)
        od.diffusion(order, phi, g, exp, 0.7, 1.3, True, (nb, nb, 0),
                     (shape[0] - 2 * nb, shape[1] - 2 * nb, shape[2]))
        out = tb.zeros(shape)
        HorizontalDiffusion.factory(name, shape, 0.7, 1.3, 0.5, 1.0, min(2, shape[2]))(dev(phi), out)
        eq(out, exp)
    for order, name in ((1, "first_order"), (2, "second_order"), (3, "third_order")):
        g = np.zeros(shape)
        g[...] = od.vertical_profile(0.03, 0.24, min(2, shape[2]), shape[2])[None, None, :]
        exp = np.zeros(shape)
        od.horizontal_smoothing(order, phi, g, exp)
        out = tb.zeros(shape)
        HorizontalSmoothing.factory(name, shape, 0.03, 0.24, min(2, shape[2]))(dev(phi), out)
        eq(out, exp)


@pytest.mark.parametrize("order", (1, 2, 3, 4, 5, 6))
def test_burgers_forward_euler_oracle(order):
    from oracle import burgers as obg

    rng = np.random.default_rng(100 + order)
    shape = (37, 29, 1)
    e = (order + 1) // 2
    u, v, ut, vt = (rng.uniform(-2, 2, size=shape) for _ in range(4))
    tu, tv = rng.uniform(-0.1, 0.1, size=shape), rng.uniform(-0.1, 0.1, size=shape)
    dt, dx, dy = 0.01, 0.011, 0.013
    origin, domain = (e, e, 0), (shape[0] - 2 * e, shape[1] - 2 * e, 1)
    for with_tnd in (False, True):
        eu, ev = np.zeros(shape), np.zeros(shape)
        obg.forward_euler(order, u, v, ut, vt, eu, ev, dt=dt, dx=dx, dy=dy, origin=origin,
                          domain=domain, u_tnd=tu if with_tnd else None,
                          v_tnd=tv if with_tnd else None)
        ou, ov = tb.zeros(shape), tb.zeros(shape)
        name = list(ADVECTION)[order - 1]
        kw = dict(in_u_tnd=dev(tu), in_v_tnd=dev(tv)) if with_tnd else {}
        stencil("forward_euler", advection=ADVECTION[name], extent=e, tnd_u=with_tnd,
                tnd_v=with_tnd)(in_u=dev(u), in_v=dev(v), in_u_tmp=dev(ut), in_v_tmp=dev(vt),
                                out_u=ou, out_v=ov, dt=dt, dx=dx, dy=dy, origin=origin,
                                domain=domain, **kw)
        eq(ou, eu)
        eq(ov, ev)


def test_periodic_enforce_oracle():
    from tasmania_b200.boundary import Periodic

    rng = np.random.default_rng(5)
    nx, ny, nz, nb = 11, 9, 3, 3
    for name, (ax, ay) in (("air_isentropic_density", (0, 0)), ("x_velocity_at_u_locations", (1, 0)),
                           ("y_velocity_at_v_locations", (0, 1))):
        phys = rng.standard_normal((nx + ax, ny + ay, nz))
        ohb = ob.Periodic(nx, ny, nz, nb)
        exp = ohb.get_numerical_field(phys, name)
        hb = Periodic(nx, ny, nz, nb)
        got = hb.get_numerical_field(phys, name)
        eq(got, exp)


def test_sliced_views_and_errors():
    """Non-contiguous views go through the explicit strides of the C ABI; host arrays and
    out-of-range boxes are refused loudly."""
    rng = np.random.default_rng(3)
    a = rng.standard_normal((12, 10, 6))
    A, O = dev(a), tb.zeros((12, 10, 6))
    stencil("copy")(src=A[:, :, 2:3], dst=O[:, :, 4:5], origin=(0, 0, 0), domain=(12, 10, 1))
    exp = np.zeros_like(a)
    exp[:, :, 4] = a[:, :, 2]
    eq(O, exp)
    with pytest.raises(tb.B200Error):
        stencil("copy")(src=a, dst=O, origin=(0, 0, 0), domain=(12, 10, 6))  # host array
    with pytest.raises(tb.B200Error):
        stencil("copy")(src=A, dst=O, origin=(0, 0, 0), domain=(13, 10, 6))  # box too large
    with pytest.raises(tb.FactoryRegistryError):
        tb.compile_stencil("no_such_stencil")
    stencil("copy")(src=A, dst=O, origin=(0, 0, 0), domain=(0, 10, 6))  # empty box is a no-op
