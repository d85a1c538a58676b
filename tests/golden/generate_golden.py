# -*- coding: utf-8 -*-
"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE's own numpy
code in place from /root/reference (see refload.py).  Run in the build container:

    python tests/golden/generate_golden.py [case ...]

Every fixture stores inputs AND reference outputs so that the tests need nothing but the
.npz file.  Fixtures are small (a few hundred kB) on purpose; the size-independent
properties are exercised at full size in the ``-m gpu`` tests.

Seeds are fixed; the reference tests' value ranges are reused (tests/conf.py:L137-L175:
s in [10, 1000], u, v in [-50, 50], q in [0, 5]).
"""
from __future__ import annotations

import os
import sys
import types
import warnings
from datetime import datetime, timedelta

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore", category=SyntaxWarning)

import refload  # noqa: E402

refload.install()
refload.install_framework()
DataArray = refload.DataArray

S, SU, SV = "air_isentropic_density", "x_momentum_isentropic", "y_momentum_isentropic"
U, V = "x_velocity_at_u_locations", "y_velocity_at_v_locations"
MTG = "montgomery_potential"
P, EXN, H = (
    "air_pressure_on_interface_levels",
    "exner_function_on_interface_levels",
    "height_on_interface_levels",
)
MFWV = "mass_fraction_of_water_vapor_in_air"
MFCW = "mass_fraction_of_cloud_liquid_water_in_air"
MFPW = "mass_fraction_of_precipitation_water_in_air"


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} kB)")


def da(x, units):
    return DataArray(x, attrs={"units": units})


# ============================================================================ 1-D dwarfs
def gen_stencils_1d():
    """K8 / K9, one-dimensional variants: the reference's SecondOrder1DX/1DY, FourthOrder1DX/1DY
    diffusers and First/Second/ThirdOrder1DX/1DY smoothers, on a row (n x 1) and a column
    (1 x n) of points and -- the classes accept it -- on a 2-D grid swept along one axis."""
    rng = np.random.default_rng(20261019)
    opts = refload.load("tasmania.framework.options")
    for name in ("second_order", "fourth_order"):
        refload.load("tasmania.dwarfs.subclasses.horizontal_diffusers." + name)
    for name in ("first_order", "second_order", "third_order"):
        refload.load("tasmania.dwarfs.subclasses.horizontal_smoothers." + name)
    hd = refload.load("tasmania.dwarfs.horizontal_diffusion")
    hsm = refload.load("tasmania.dwarfs.horizontal_smoothing")
    dx, dy = 1100.0, 900.0
    out = {"scalars": np.array([dx, dy])}
    shapes = {"x": {"row": (19, 1, 5), "grid": (14, 6, 4)}, "y": {"row": (1, 21, 5), "grid": (7, 13, 4)}}
    kw = dict(backend="numpy", backend_options=opts.BackendOptions(),
              storage_options=opts.StorageOptions())
    for ax in ("x", "y"):
        for tag, shape in shapes[ax].items():
            phi = rng.uniform(-5, 5, size=shape)
            base = rng.uniform(-5, 5, size=shape)
            out[f"{ax}_{tag}_phi"], out[f"{ax}_{tag}_base"] = phi, base
            for order, name in ((2, "second_order"), (4, "fourth_order")):
                obj = hd.HorizontalDiffusion.factory(f"{name}_1d{ax}", shape, dx, dy, 0.5, 1.0, 3, **kw)
                tnd = np.zeros(shape)
                obj(phi, tnd, overwrite_output=True)
                acc = base.copy()
                obj(phi, acc, overwrite_output=False)
                out[f"k8_{order}_{ax}_{tag}_tnd"], out[f"k8_{order}_{ax}_{tag}_acc"] = tnd, acc
            for order, name in ((1, "first_order"), (2, "second_order"), (3, "third_order")):
                obj = hsm.HorizontalSmoothing.factory(f"{name}_1d{ax}", shape, 0.03, 0.24, 3, **kw)
                sm = np.zeros(shape)
                obj(phi, sm)
                out[f"k9_{order}_{ax}_{tag}_out"] = sm
    # ---- the global `thomas` stencil (stencil_definitions/cla.py): diagonally dominant random
    # systems, plus columns engineered to hit the zero-pivot branch (eliminated diagonal == 0
    # at level 1: b = 2, a = 1, beta[0] = 1, c[0] = 2)
    cla = refload.load("tasmania.framework.subclasses.stencil_definitions.cla")
    shape = (6, 5, 9)
    a, c, d = (rng.uniform(-1, 1, size=shape) for _ in range(3))
    b = rng.uniform(2.5, 4, size=shape) * rng.choice([-1.0, 1.0], size=shape)
    for (i, j) in ((1, 1), (3, 2), (4, 4)):
        b[i, j, 1], c[i, j, 1] = 1.0, 2.0
        a[i, j, 2], b[i, j, 2] = 1.0, 2.0
    x = np.zeros(shape)
    origin, domain = (1, 0, 1), (5, 5, 7)
    with np.errstate(divide="ignore", invalid="ignore"):
        cla.thomas_numpy(a, b, c, d, x, origin=origin, domain=domain)
    assert np.isfinite(x).all()
    out.update(thomas_a=a, thomas_b=b, thomas_c=c, thomas_d=d, thomas_x=x,
               thomas_box=np.array(origin + domain))
    # ---- the class-less `diffusion` stencil (hyperdiffusion filter, stencil_definitions/diffusion.py)
    dif = refload.load("tasmania.framework.subclasses.stencil_definitions.diffusion")
    shape = (15, 13, 5)
    hphi = rng.uniform(-5, 5, size=shape)
    hout = np.zeros(shape)
    dif.diffusion_numpy(hphi, hout, alpha=1.0 / 32.0, origin=(3, 3, 1), domain=(9, 7, 3))
    out.update(hyper_phi=hphi, hyper_out=hout, hyper_box=np.array([3, 3, 1, 9, 7, 3]),
               hyper_alpha=np.array(1.0 / 32.0))
    # ---- the 1-D lateral boundaries (Relaxed1DX / 1DY, Periodic1DX / 1DY of the reference)
    bnames = ("air_isentropic_density", "x_velocity_at_u_locations", "y_velocity_at_v_locations",
              "air_pressure_on_interface_levels")
    for kind, kw, nb in (("relaxed", {"nr": 5}, 2), ("periodic", {}, 2)):
        for ax, (bnx, bny) in (("x", (17, 1)), ("y", (1, 15))):
            bnz = 4
            hb = _make_domain(bnx, bny, bnz, kind, nb, kw, topo=False).horizontal_boundary
            bshape = (hb.ni + 1, hb.nj + 1, bnz + 1)
            tag = f"hb_{kind}_{ax}"
            out[tag + "_dims"] = np.array([bnx, bny, bnz, nb, kw.get("nr", 0)])
            phys = rng.standard_normal((bnx, bny, bnz))
            out[tag + "_phys"] = phys
            out[tag + "_num"] = np.asarray(hb.get_numerical_field(phys.copy(), field_name=bnames[0]))
            refs = {n: rng.standard_normal(bshape) for n in bnames}
            hb.reference_state = {n: DataArray(v.copy(), attrs={"units": "1"}) for n, v in refs.items()}
            for m, n in enumerate(bnames):
                f = rng.standard_normal(bshape)
                out[f"{tag}_ref{m}"], out[f"{tag}_in{m}"] = refs[n], f.copy()
                hb.enforce_field(f, field_name=n, field_units="1")
                hb.set_outermost_layers_x(f, field_name=n, field_units="1")
                hb.set_outermost_layers_y(f, field_name=n, field_units="1")
                out[f"{tag}_out{m}"] = f
    save("stencils_1d", **out)


# ============================================================================ stencils
def gen_stencils():
    """Per-stencil fixtures: K1, K2 (4 flux schemes), K3, K4, K5, K6, K7, K8, K9, K12."""
    rng = np.random.default_rng(20261018)
    nx, ny, nz = 17, 15, 6
    shape = (nx + 1, ny + 1, nz + 1)
    out = {}

    def field(lo, hi):
        return rng.uniform(lo, hi, size=shape)

    s_now, s_int = field(10, 1000), field(10, 1000)
    u_int, v_int = field(-50, 50), field(-50, 50)
    su_now, su_int = field(-5e3, 5e3), field(-5e3, 5e3)
    sv_now, sv_int = field(-5e3, 5e3), field(-5e3, 5e3)
    mtg_now, mtg_new = field(2.9e5, 3.1e5), field(2.9e5, 3.1e5)
    s_tnd, su_tnd, sv_tnd = field(-1, 1), field(-10, 10), field(-10, 10)
    sq_now = [field(0, 5) for _ in range(3)]
    sq_int = [field(0, 5) for _ in range(3)]
    q_tnd = [field(-1e-3, 1e-3) for _ in range(3)]
    s_new_in = field(10, 1000)
    dt, dx, dy, eps = 1.5, 1100.0, 900.0, 0.3
    out.update(
        s_now=s_now, s_int=s_int, u_int=u_int, v_int=v_int, su_now=su_now, su_int=su_int,
        sv_now=sv_now, sv_int=sv_int, mtg_now=mtg_now, mtg_new=mtg_new, s_tnd=s_tnd,
        su_tnd=su_tnd, sv_tnd=sv_tnd, s_new_in=s_new_in,
        sq_now=np.stack(sq_now), sq_int=np.stack(sq_int), q_tnd=np.stack(q_tnd),
        scalars=np.array([dt, dx, dy, eps]),
    )

    utils = refload.load("tasmania.isentropic.dynamics.subclasses.prognostics.utils")
    base = "tasmania.isentropic.dynamics.subclasses.minimal_horizontal_fluxes."
    classes = {
        "upwind": "Upwind",
        "centered": "Centered",
        "third_order_upwind": "ThirdOrderUpwind",
        "fifth_order_upwind": "FifthOrderUpwind",
    }
    for scheme, cls in classes.items():
        mod = refload.load(base + scheme)
        hflux = getattr(mod, cls)(backend="numpy")
        e = hflux.extent
        externals = dict(hflux.externals or {})
        if scheme == "centered":
            # the numpy Centered flux looks its helpers up as globals
            full = refload.load(
                "tasmania.isentropic.dynamics.subclasses.horizontal_fluxes.centered"
            )
            mod.get_centered_flux_x = full.Centered.get_centered_flux_x_numpy
            mod.get_centered_flux_y = full.Centered.get_centered_flux_y_numpy
        for moist in (False, True):
            for tnd in (False, True):
                ext = dict(externals)
                ext.update(
                    extent=e, moist=moist,
                    flux_dry=hflux.get_subroutine_definition("flux_dry"),
                    flux_moist=hflux.get_subroutine_definition("flux_moist"),
                )
                k1 = refload.numpy_stencil(utils.step_forward_euler_numpy, ext)
                k2 = refload.numpy_stencil(utils.step_forward_euler_momentum_numpy, ext)
                nb = e
                origin, domain = (nb, nb, 0), (nx - 2 * nb, ny - 2 * nb, nz)
                s_new = np.zeros(shape)
                sq_new = [np.zeros(shape) for _ in range(3)]
                kw = {}
                if moist:
                    kw = dict(
                        sqv_now=sq_now[0], sqv_int=sq_int[0], sqv_new=sq_new[0],
                        sqc_now=sq_now[1], sqc_int=sq_int[1], sqc_new=sq_new[1],
                        sqr_now=sq_now[2], sqr_int=sq_int[2], sqr_new=sq_new[2],
                    )
                    if tnd:
                        kw.update(qv_tnd=q_tnd[0], qc_tnd=q_tnd[1], qr_tnd=q_tnd[2])
                k1(
                    s_now=s_now, s_int=s_int, s_new=s_new, u_int=u_int, v_int=v_int,
                    su_int=su_int, sv_int=sv_int, s_tnd=s_tnd if tnd else None,
                    dt=dt, dx=dx, dy=dy, origin=origin, domain=domain, **kw,
                )
                tag = f"{scheme}_m{int(moist)}_t{int(tnd)}"
                out[f"k1_{tag}_s_new"] = s_new
                if moist:
                    out[f"k1_{tag}_sq_new"] = np.stack(sq_new)
                if not moist:
                    su_new, sv_new = np.zeros(shape), np.zeros(shape)
                    k2(
                        s_now=s_now, s_int=s_int, s_new=s_new_in, u_int=u_int, v_int=v_int,
                        su_now=su_now, su_int=su_int, su_new=su_new, sv_now=sv_now,
                        sv_int=sv_int, sv_new=sv_new, mtg_now=mtg_now, mtg_new=mtg_new,
                        su_tnd=su_tnd if tnd else None, sv_tnd=sv_tnd if tnd else None,
                        dt=dt, dx=dx, dy=dy, eps=eps, origin=origin, domain=domain,
                    )
                    out[f"k2_{tag}_su_new"] = su_new
                    out[f"k2_{tag}_sv_new"] = sv_new

    # ---- K3 diagnostics
    diag = refload.load("tasmania.isentropic.dynamics.diagnostics")
    consts = {"pref": 1.0e5, "rd": 287.05, "g": 9.80665, "cp": 1004.0}
    theta1d = np.linspace(400.0, 280.0, nz + 1)
    theta = np.zeros(shape)
    theta[:nx, :ny, :] = theta1d[None, None, :]
    hs = np.zeros(shape)
    hs[:nx, :ny, nz] = rng.uniform(0, 800, size=(nx, ny))
    s_col = rng.uniform(5, 60, size=shape)
    dz, pt = 20.0, 11868.9
    D = diag.IsentropicDiagnostics
    mont = refload.numpy_stencil(D._montgomery_numpy, consts)
    dvar = refload.numpy_stencil(D._diagnostic_variables_numpy, consts)
    hgt = refload.numpy_stencil(D._height_numpy, consts)
    dat = refload.numpy_stencil(D._density_and_temperature_numpy, consts)
    mtg = np.zeros(shape)
    mont(in_hs=hs, in_s=s_col, inout_mtg=mtg, dz=dz, pt=pt, theta_s=theta1d[-1],
         origin=(0, 0, 0), domain=(nx, ny, nz + 1))
    p, exn, mtg2, h = (np.zeros(shape) for _ in range(4))
    dvar(in_theta=theta, in_hs=hs, in_s=s_col, inout_p=p, out_exn=exn, inout_mtg=mtg2,
         inout_h=h, dz=dz, pt=pt, origin=(0, 0, 0), domain=(nx, ny, nz + 1))
    h2 = np.zeros(shape)
    hgt(in_theta=theta, in_hs=hs, in_s=s_col, inout_h=h2, dz=dz, pt=pt,
        origin=(0, 0, 0), domain=(nx, ny, nz + 1))
    rho, temp = np.zeros(shape), np.zeros(shape)
    dat(in_theta=theta, in_s=s_col, in_exn=exn, in_h=h, out_rho=rho, out_t=temp,
        origin=(0, 0, 0), domain=(nx, ny, nz))
    out.update(k3_theta=theta, k3_hs=hs, k3_s=s_col, k3_scalars=np.array([dz, pt, theta1d[-1]]),
               k3_mtg=mtg, k3_p=p, k3_exn=exn, k3_mtg2=mtg2, k3_h=h, k3_h2=h2, k3_rho=rho,
               k3_t=temp)

    # ---- K4 velocity / momenta, K7 density / mass fraction
    dd = refload.load("tasmania.dwarfs.diagnostics")
    HV, WC = dd.HorizontalVelocity, dd.WaterConstituent
    vx = refload.numpy_stencil(HV._diagnose_velocity_x_numpy, {"staggering": True})
    vy = refload.numpy_stencil(HV._diagnose_velocity_y_numpy, {"staggering": True})
    mom = refload.numpy_stencil(HV._diagnose_momenta_numpy, {"staggering": True})
    u_out, v_out = np.zeros(shape), np.zeros(shape)
    vx(in_d=s_now, in_du=su_now, out_u=u_out, origin=(1, 0, 0), domain=(nx - 1, ny, nz))
    vy(in_d=s_now, in_dv=sv_now, out_v=v_out, origin=(0, 1, 0), domain=(nx, ny - 1, nz))
    du_out, dv_out = np.zeros(shape), np.zeros(shape)
    mom(in_d=s_now, in_u=u_int, in_v=v_int, out_du=du_out, out_dv=dv_out,
        origin=(0, 0, 0), domain=(nx, ny, nz))
    q = rng.uniform(-0.5, 5, size=shape)
    dens = refload.numpy_stencil(WC._diagnose_density_numpy, {"clipping": True})
    mf = refload.numpy_stencil(WC._diagnose_mass_fraction_numpy, {"clipping": True})
    sq_o, q_o = np.zeros(shape), np.zeros(shape)
    dens(in_d=s_now, in_q=q, out_dq=sq_o, origin=(0, 0, 0), domain=(nx, ny, nz))
    sq_in = rng.uniform(-100, 3000, size=shape)
    mf(in_d=s_now, in_dq=sq_in, out_q=q_o, origin=(0, 0, 0), domain=(nx, ny, nz))
    out.update(k4_u=u_out, k4_v=v_out, k4_du=du_out, k4_dv=dv_out, k7_q=q, k7_sq=sq_o,
               k7_sq_in=sq_in, k7_q_out=q_o)

    # ---- K5 irelax / relax with the real Relaxed gamma
    dom = _make_domain(nx, ny, nz, "relaxed", 3, {"nr": 6})
    hb = dom.horizontal_boundary
    gamma = np.array(hb._gamma)
    alg = refload.load("tasmania.framework.subclasses.stencil_definitions.algorithms")
    phi, phi_ref = field(10, 1000), field(10, 1000)
    phi_io = phi.copy()
    alg.irelax_numpy(gamma, phi_ref, phi_io, origin=(0, 0, 0), domain=(nx, ny, nz))
    phi_o = np.zeros(shape)
    alg.relax_numpy(gamma, phi, phi_ref, phi_o, origin=(0, 0, 0), domain=(nx + 1, ny, nz))
    out.update(k5_gamma=gamma, k5_phi=phi, k5_phi_ref=phi_ref, k5_irelax=phi_io, k5_relax=phi_o)

    # ---- K6 Rayleigh damping (real coefficient matrix)
    refload.load("tasmania.dwarfs.subclasses.vertical_dampers.rayleigh")
    vd = refload.load("tasmania.dwarfs.vertical_damping")
    opts = refload.load("tasmania.framework.options")
    g = dom.numerical_grid
    damper = vd.VerticalDamping.factory(
        "rayleigh", g, 4, 5e-4, backend="numpy", backend_options=opts.BackendOptions(),
        storage_shape=shape, storage_options=opts.StorageOptions(),
    )
    phi_out = np.zeros(shape)
    damper(timedelta(seconds=7), phi, phi_ref, s_now, phi_out)
    out.update(k6_rmat=np.array(damper._rmat), k6_out=phi_out, k6_z=g.z.values,
               k6_zhl=g.z_on_interface_levels.values, k6_params=np.array([4, 5e-4, 7.0]))

    # ---- K8 diffusion, K9 smoothing
    refload.load("tasmania.dwarfs.subclasses.horizontal_diffusers.second_order")
    refload.load("tasmania.dwarfs.subclasses.horizontal_diffusers.fourth_order")
    hd = refload.load("tasmania.dwarfs.horizontal_diffusion")
    for order, name in ((2, "second_order"), (4, "fourth_order")):
        obj = hd.HorizontalDiffusion.factory(
            name, shape, dx, dy, 0.5, 1.0, 3, backend="numpy",
            backend_options=opts.BackendOptions(), storage_options=opts.StorageOptions(),
        )
        tnd = np.zeros(shape)
        obj(phi, tnd, overwrite_output=True)
        acc = phi_ref.copy()
        obj(phi, acc, overwrite_output=False)
        out[f"k8_{order}_gamma"] = np.array(obj._gamma)
        out[f"k8_{order}_tnd"] = tnd
        out[f"k8_{order}_acc"] = acc
    hsm = refload.load("tasmania.dwarfs.horizontal_smoothing")
    for order, name in ((1, "first_order"), (2, "second_order"), (3, "third_order")):
        refload.load("tasmania.dwarfs.subclasses.horizontal_smoothers." + name)
        obj = hsm.HorizontalSmoothing.factory(
            name, shape, 0.03, 0.24, 3, backend="numpy",
            backend_options=opts.BackendOptions(), storage_options=opts.StorageOptions(),
        )
        sm = np.zeros(shape)
        obj(phi, sm)
        out[f"k9_{order}_gamma"] = np.array(obj._gamma)
        out[f"k9_{order}_out"] = sm

    # ---- K12 elementwise
    m = refload.load("tasmania.framework.subclasses.stencil_definitions.math")
    cp = refload.load("tasmania.framework.subclasses.stencil_definitions.copy")
    a, b, c = field(-5, 5), field(-5, 5), field(-5, 5)
    box = dict(origin=(1, 2, 0), domain=(nx - 2, ny - 3, nz))
    res = {}

    def run(name, fn, *ins, **kw):
        o = np.zeros(shape)
        fn(*ins, o, **kw, **box)
        res[name] = o

    run("abs", m.abs_numpy, a)
    run("add", m.add_numpy, a, b)
    run("addsub", m.addsub_numpy, a, b, c)
    run("clip", m.clip_numpy, a)
    run("fma", m.fma_numpy, a, b, f=0.37)
    run("mul", m.mul_numpy, a, b)
    run("scale", m.scale_numpy, a, f=-1.7)
    run("sub", m.sub_numpy, a, b)
    run("copy", cp.copy_numpy, a)
    run("copychange", cp.copychange_numpy, a)
    run("sts_rk2_0", alg.sts_rk2_0_numpy, a, b, c, dt=0.8)
    run("sts_rk3ws_0", alg.sts_rk3ws_0_numpy, a, b, c, dt=0.8)
    for name, fn, ins, kw in (
        ("iabs", m.iabs_numpy, (), {}),
        ("iadd", m.iadd_numpy, (b,), {}),
        ("iaddsub", m.iaddsub_numpy, (b, c), {}),
        ("iclip", m.iclip_numpy, (), {}),
        ("imul", m.imul_numpy, (b,), {}),
        ("iscale", m.iscale_numpy, (), {"f": 2.5}),
        ("isub", m.isub_numpy, (b,), {}),
    ):
        io = a.copy()
        fn(io, *ins, **kw, **box)
        res[name] = io
    out.update(k12_a=a, k12_b=b, k12_c=c, **{f"k12_{k}": v for k, v in res.items()})

    save("stencils", dims=np.array([nx, ny, nz]), **out)



# ============================================================================ Kessler (K11)
def gen_kessler():
    """K11: kessler, saturation (diagnostic + prognostic), fall_velocity, sedimentation
    (first / second order upwind flux), accumulated_precipitation -- the reference's numpy
    stencil definitions of src/tasmania/physics/microphysics/{kessler,utils}.py with the
    externals their classes inject (kessler.py:L183-L190, L554-L562, L1283-L1287,
    utils.py:L205-L207)."""
    rng = np.random.default_rng(20261019)
    nx, ny, nz = 13, 11, 8
    shape = (nx + 1, ny + 1, nz + 1)
    out = {}
    rd, rv, cp, lhvw, pref, rhow = 287.05, 461.52, 1004.0, 2.5e6, 1.0e5, 1.0e3
    beta = rd / rv

    # a physically plausible column: pressure increasing with k, height decreasing
    p_hl = np.zeros(shape)
    p_hl[...] = np.linspace(1.5e4, 1.0e5, nz + 1)[None, None, :] * rng.uniform(0.97, 1.03, size=shape)
    exn_hl = cp * (p_hl / pref) ** (rd / cp)
    p_ml, exn_ml = rng.uniform(2e4, 9.5e4, size=shape), rng.uniform(600, 1000, size=shape)
    t = rng.uniform(215.0, 305.0, size=shape)
    rho = rng.uniform(0.15, 1.3, size=shape)
    qv = rng.uniform(0.0, 0.02, size=shape)
    qc = rng.uniform(0.0, 3e-3, size=shape)
    qr = rng.uniform(-5e-4, 3e-3, size=shape)   # negatives and zeros exercise the where() branches
    qr[::3, ::2, ::2] = 0.0
    h_hl = np.zeros(shape)
    h_hl[...] = np.linspace(1.6e4, 0.0, nz + 1)[None, None, :] + rng.uniform(-150, 150, size=shape)
    vt_in = rng.uniform(0.0, 9.0, size=shape)
    accprec = rng.uniform(0.0, 4.0, size=(nx + 1, ny + 1, 1))
    prev = {n: rng.uniform(-1e-5, 1e-5, size=shape) for n in ("qc", "qr", "qv", "theta")}
    a, k1, k2, dt, sr = 1.0e-3, 1.0e-3, 2.2, 7.5, 0.025
    out.update(p_hl=p_hl, exn_hl=exn_hl, p_ml=p_ml, exn_ml=exn_ml, t=t, rho=rho, qv=qv, qc=qc,
               qr=qr, h_hl=h_hl, vt_in=vt_in, accprec=accprec,
               prev_qc=prev["qc"], prev_qr=prev["qr"], prev_qv=prev["qv"], prev_theta=prev["theta"],
               scalars=np.array([a, k1, k2, dt, sr, rd, rv, cp, lhvw, rhow]))

    gen = refload.load("tasmania.framework.subclasses.subroutine_definitions.generics")
    ke = refload.load("tasmania.physics.microphysics.kessler")
    ut = refload.load("tasmania.physics.microphysics.utils")
    box = dict(origin=(0, 0, 0), domain=(nx, ny, nz))

    # ---- kessler
    for apoil in (True, False):
        for evap in (True, False):
            for ow in (True, False):
                ext = {"air_pressure_on_interface_levels": apoil, "beta": beta, "e": np.exp(1),
                       "lhvw": lhvw, "rain_evaporation": evap, "set_output": gen.set_output_numpy}
                st = refload.numpy_stencil(ke.KesslerMicrophysics._kessler_numpy, ext)
                o = {n: (np.zeros(shape) if ow else prev[n].copy()) for n in prev}
                st(in_rho=rho, in_p=p_hl if apoil else p_ml, in_t=t,
                   in_exn=exn_hl if apoil else exn_ml, in_qc=qc, in_qr=qr, in_qv=qv,
                   out_qc_tnd=o["qc"], out_qr_tnd=o["qr"], out_qv_tnd=o["qv"],
                   out_theta_tnd=o["theta"], a=a, k1=k1, k2=k2, ow_out_qc_tnd=ow,
                   ow_out_qr_tnd=ow, ow_out_qv_tnd=ow, ow_out_theta_tnd=ow, **box)
                tag = f"kessler_p{int(apoil)}_e{int(evap)}_o{int(ow)}"
                for n in o:
                    out[f"{tag}_{n}"] = o[n]

    # ---- saturation adjustment
    for apoil in (True, False):
        ext = {"air_pressure_on_interface_levels": apoil, "beta": beta, "cp": cp, "e": np.exp(1),
               "lhvw": lhvw, "rv": rv, "set_output": gen.set_output_numpy}
        pp, ee = (p_hl, exn_hl) if apoil else (p_ml, exn_ml)
        for ow in (True, False):
            st = refload.numpy_stencil(ke.KesslerSaturationAdjustmentDiagnostic._saturation_diagnostic_numpy, ext)
            o_qv, o_qc, o_t = np.zeros(shape), np.zeros(shape), np.zeros(shape)
            tnd = np.zeros(shape) if ow else prev["theta"].copy()
            st(in_p=pp, in_t=t, in_exn=ee, in_qv=qv, in_qc=qc, out_qv=o_qv, out_qc=o_qc, out_t=o_t,
               tnd_theta=tnd, dt=dt, ow_tnd_theta=ow, **box)
            tag = f"satd_p{int(apoil)}_o{int(ow)}"
            out.update({f"{tag}_qv": o_qv, f"{tag}_qc": o_qc, f"{tag}_t": o_t, f"{tag}_theta": tnd})
            st = refload.numpy_stencil(ke.KesslerSaturationAdjustmentPrognostic._saturation_prognostic_numpy, ext)
            o = {n: (np.zeros(shape) if ow else prev[n].copy()) for n in ("qv", "qc", "theta")}
            st(in_p=pp, in_t=t, in_exn=ee, in_qv=qv, in_qc=qc, tnd_qv=o["qv"], tnd_qc=o["qc"],
               tnd_theta=o["theta"], sr=sr, ow_tnd_qv=ow, ow_tnd_qc=ow, ow_tnd_theta=ow, **box)
            tag = f"satp_p{int(apoil)}_o{int(ow)}"
            out.update({f"{tag}_{n}": o[n] for n in o})

    # ---- fall velocity (array_call copies the surface density slab first, kessler.py:L1169)
    rho_s = np.zeros(shape)
    rho_s[:nx, :ny, :nz] = rho[:nx, :ny, nz - 1 : nz]
    vt = np.zeros(shape)
    ke.KesslerFallVelocity._fall_velocity_numpy(in_rho=rho, in_rho_s=rho_s, in_qr=qr, out_vt=vt, **box)
    out.update(fall_rho_s=rho_s, fall_vt=vt)

    # ---- sedimentation
    for order, mod, cls in ((1, "first_order", "FirstOrderUpwind"), (2, "second_order", "SecondOrderUpwind")):
        sf = getattr(refload.load("tasmania.physics.microphysics.sedimentation_fluxes." + mod), cls)
        for ow in (True, False):
            ext = {"set_output": gen.set_output_numpy, "sflux": sf.call_numpy, "sflux_extent": sf.nb}
            st = refload.numpy_stencil(ke.KesslerSedimentation._sedimentation_numpy, ext)
            tnd = rng.uniform(-1, 1, size=shape) if ow else prev["qr"].copy()
            st(in_rho=rho, in_h=h_hl, in_qr=qr, in_vt=vt_in, out_tnd_qr=tnd, ow_out_tnd_qr=ow, **box)
            out[f"sed_{order}_o{int(ow)}"] = tnd

    # ---- accumulated precipitation on the surface slabs (utils.py:L262-L281)
    st = refload.numpy_stencil(ut.Precipitation._accumulated_precipitation_numpy, {"rhow": rhow})
    prec, acc = np.zeros((nx + 1, ny + 1, 1)), np.zeros((nx + 1, ny + 1, 1))
    st(in_rho=rho[:, :, nz - 1 : nz], in_qr=qr[:, :, nz - 1 : nz], in_vt=vt_in[:, :, nz - 1 : nz],
       in_accprec=accprec[:, :, :1], out_prec=prec[:, :, :1], out_accprec=acc[:, :, :1], dt=dt,
       origin=(0, 0, 0), domain=(nx, ny, 1))
    out.update(prec=prec, acc=acc)

    save("kessler", dims=np.array([nx, ny, nz]), **out)

# ============================================================================ isentropic physics (8f)
VFLUX_CLASSES = {"upwind": "Upwind", "centered": "Centered", "third_order_upwind": "ThirdOrderUpwind",
                 "fifth_order_upwind": "FifthOrderUpwind"}


def gen_vertical_advection():
    """IsentropicVerticalAdvection._stencil_numpy
    (src/tasmania/isentropic/physics/vertical_advection.py:L271-L386) with the four minimal
    vertical flux schemes (src/tasmania/isentropic/dynamics/subclasses/minimal_vertical_fluxes/
    *.py), vertical velocity on main or interface levels, dry and moist, overwrite on / off.
    The method is run unbound on a stand-in ``self`` that carries what it reads: _vflux, _stgz,
    _moist, get_field_storage_shape, storage_options.dtype."""
    va = refload.load("tasmania.isentropic.physics.vertical_advection")
    gen = refload.load("tasmania.framework.subclasses.subroutine_definitions.generics")
    st = refload.numpy_stencil(va.IsentropicVerticalAdvection._stencil_numpy,
                               {"set_output": gen.set_output_numpy})
    rng = np.random.default_rng(20261020)
    nx, ny, nz = 11, 9, 13
    shape = (nx + 1, ny + 1, nz + 1)
    dz = 2.5
    ins = {
        "w": rng.uniform(-0.02, 0.02, size=shape),
        "s": rng.uniform(10.0, 1000.0, size=shape),
        "su": rng.uniform(-5e4, 5e4, size=shape),
        "sv": rng.uniform(-5e4, 5e4, size=shape),
        "qv": rng.uniform(0.0, 5.0, size=shape),
        "qc": rng.uniform(0.0, 5.0, size=shape),
        "qr": rng.uniform(0.0, 5.0, size=shape),
    }
    ins["w"][::4, ::3, ::2] = 0.0  # zeros exercise the upwind selection
    prev = {n: rng.uniform(-1.0, 1.0, size=shape) for n in ("s", "su", "sv", "qv", "qc", "qr")}
    out = {"in_" + k: v for k, v in ins.items()}
    out.update({"prev_" + k: v for k, v in prev.items()})
    for scheme, cls in VFLUX_CLASSES.items():
        mod = refload.load("tasmania.isentropic.dynamics.subclasses.minimal_vertical_fluxes." + scheme)
        flux_cls = getattr(mod, cls)
        vflux = types.SimpleNamespace(
            extent=flux_cls.extent,
            get_subroutine_definition=lambda name, c=flux_cls: getattr(c, name + "_numpy"))
        for stgz in (False, True):
            for moist in (False, True):
                for ow in (True, False):
                    fake = types.SimpleNamespace(
                        _vflux=vflux, _stgz=stgz, _moist=moist,
                        get_field_storage_shape=lambda name=None: shape,
                        storage_options=types.SimpleNamespace(dtype=np.float64))
                    names = ("s", "su", "sv") + (("qv", "qc", "qr") if moist else ())
                    outs = {n: (rng.uniform(-1, 1, size=shape) if ow else prev[n].copy()) for n in names}
                    kw = {"in_w": ins["w"], "dz": dz, "origin": (0, 0, 0), "domain": (nx, ny, nz)}
                    for n in names:
                        kw["in_" + n] = ins[n]
                        kw["out_" + n] = outs[n]
                        kw["ow_out_" + n] = ow
                    st(fake, **kw)
                    for n in names:
                        out[f"{scheme}_z{int(stgz)}_m{int(moist)}_o{int(ow)}_{n}"] = outs[n]
    # ---- Coriolis: IsentropicConservativeCoriolis._stencil_numpy,
    # src/tasmania/isentropic/physics/coriolis.py:L166-L186, on the interior box (nb = 2)
    co = refload.load("tasmania.isentropic.physics.coriolis")
    cst = refload.numpy_stencil(co.IsentropicConservativeCoriolis._stencil_numpy,
                                {"set_output": gen.set_output_numpy})
    f = 1.2345e-4
    for owu, owv in ((True, True), (False, True), (False, False)):
        tu, tv = prev["su"].copy(), prev["sv"].copy()
        cst(in_su=ins["su"], in_sv=ins["sv"], tnd_su=tu, tnd_sv=tv, f=f, ow_tnd_su=owu, ow_tnd_sv=owv,
            origin=(2, 2, 0), domain=(nx - 4, ny - 4, nz))
        out[f"coriolis_o{int(owu)}{int(owv)}_su"] = tu
        out[f"coriolis_o{int(owu)}{int(owv)}_sv"] = tv
    # ---- Smagorinsky: Smagorinsky2d / IsentropicSmagorinsky _stencil_numpy with
    # the class's own "core" subroutine (physics/turbulence.py:L165-L229,
    # isentropic/physics/turbulence.py:L99-L125), interior box of nb = 2 and nb = 3
    tu2 = refload.load("tasmania.physics.turbulence")
    tui = refload.load("tasmania.isentropic.physics.turbulence")
    ext = {"set_output": gen.set_output_numpy, "core": tu2.Smagorinsky2d._core_numpy}
    st2 = refload.numpy_stencil(tu2.Smagorinsky2d._stencil_numpy, ext)
    sti = refload.numpy_stencil(tui.IsentropicSmagorinsky._stencil_numpy, ext)
    vel = {"u": rng.uniform(-50, 50, size=shape), "v": rng.uniform(-50, 50, size=shape)}
    out.update({"in_u": vel["u"], "in_v": vel["v"]})
    sdx, sdy, cs = 1100.0, 950.0, 0.18
    for nb, ow in ((2, True), (3, False)):
        box = dict(origin=(nb, nb, 0), domain=(nx - 2 * nb, ny - 2 * nb, nz))
        a, b = prev["su"].copy(), prev["sv"].copy()
        st2(in_u=vel["u"], in_v=vel["v"], out_u_tnd=a, out_v_tnd=b, dx=sdx, dy=sdy, cs=cs,
            ow_out_u_tnd=ow, ow_out_v_tnd=not ow, **box)
        out[f"smag2d_nb{nb}_u"], out[f"smag2d_nb{nb}_v"] = a, b
        a, b = prev["su"].copy(), prev["sv"].copy()
        sti(in_s=ins["s"], in_su=ins["su"], in_sv=ins["sv"], out_su_tnd=a, out_sv_tnd=b, dx=sdx, dy=sdy,
            cs=cs, ow_out_su_tnd=ow, ow_out_sv_tnd=not ow, **box)
        out[f"smagisen_nb{nb}_su"], out[f"smagisen_nb{nb}_sv"] = a, b
    # ---- implicit vertical advection:
    # IsentropicImplicitVerticalAdvectionDiagnostic._stencil_numpy
    # (isentropic/physics/implicit_vertical_advection.py:L221-L336) with the reference's own
    # setup_thomas / thomas subroutines (framework/subclasses/subroutine_definitions/cla.py)
    iva = refload.load("tasmania.isentropic.physics.implicit_vertical_advection")
    cla = refload.load("tasmania.framework.subclasses.subroutine_definitions.cla")
    subs = {"setup_thomas": cla.setup_tridiagonal_system_numpy, "thomas": cla.thomas_numpy}
    gamma = 7.5 / (4.0 * dz)
    wbig = ins["w"] * 8.0  # |gamma w| up to 0.12: a diagonally dominant but non-trivial system
    out["in_w_implicit"] = wbig
    for stgz in (False, True):
        for moist in (False, True):
            ist = refload.numpy_stencil(iva.IsentropicImplicitVerticalAdvectionDiagnostic._stencil_numpy,
                                        {"staggering": stgz, "moist": moist})
            fake = types.SimpleNamespace(
                zeros=lambda backend=None, *, shape, storage_options=None: np.zeros(shape),
                ones=lambda backend=None, *, shape, storage_options=None: np.ones(shape),
                get_subroutine_definition=lambda name: subs[name])
            names = ("s", "su", "sv") + (("qv", "qc", "qr") if moist else ())
            outs = {n: prev[n].copy() for n in names}
            kw = {"in_w": wbig, "gamma": gamma, "origin": (0, 0, 0), "domain": (nx, ny, nz)}
            for n in names:
                kw["in_" + n], kw["out_" + n] = ins[n], outs[n]
            ist(fake, **kw)
            for n in names:
                out[f"implicit_z{int(stgz)}_m{int(moist)}_{n}"] = outs[n]
    # the prognostic variant (L793-L919): same solves, returned as tendencies
    for stgz, moist in ((False, True), (True, False)):
        pst = refload.numpy_stencil(iva.IsentropicImplicitVerticalAdvectionPrognostic._stencil_numpy,
                                    {"vstaggering": stgz, "moist": moist})
        fake = types.SimpleNamespace(
            zeros=lambda backend=None, *, shape, storage_options=None: np.zeros(shape),
            ones=lambda backend=None, *, shape, storage_options=None: np.ones(shape),
            get_subroutine_definition=lambda name: subs[name])
        names = ("s", "su", "sv") + (("qv", "qc", "qr") if moist else ())
        outs = {n: prev[n].copy() for n in names}
        kw = {"in_w": wbig, "dt": 7.5, "gamma": gamma, "origin": (0, 0, 0), "domain": (nx, ny, nz)}
        for n in names:
            kw["in_" + n], kw["tnd_" + n] = ins[n], outs[n]
        pst(fake, **kw)
        for n in names:
            out[f"implicit_tnd_z{int(stgz)}_m{int(moist)}_{n}"] = outs[n]
    save("isentropic_physics", dims=np.array([nx, ny, nz]), dz=np.array([dz]), f=np.array([f]),
         smag=np.array([sdx, sdy, cs]), gamma=np.array([gamma]), **out)


# ============================================================================ isentropic
def _make_domain(nx, ny, nz, hb_type, nb, hb_kwargs, topo_time=1800.0, xlim=(-176, 176),
                 topo=True):
    dom = refload.load("tasmania.domain.domain")
    refload.load("tasmania.domain.subclasses.horizontal_boundaries.relaxed")
    refload.load("tasmania.domain.subclasses.horizontal_boundaries.periodic")
    refload.load("tasmania.domain.subclasses.horizontal_boundaries.dirichlet")
    refload.load("tasmania.domain.subclasses.topographies.gaussian")
    refload.load("tasmania.domain.subclasses.topographies.flat")
    kw = dict(
        topography_type="gaussian",
        topography_kwargs={
            "time": timedelta(seconds=topo_time),
            "max_height": da(0.5, "km"),
            "width_x": da(50.0, "km"),
            "width_y": da(50.0, "km"),
            "smooth": False,
        },
    ) if topo else {}
    return dom.Domain(
        DataArray(list(xlim), dims="x", attrs={"units": "km"}), nx,
        DataArray(list(xlim), dims="y", attrs={"units": "km"}), ny,
        DataArray([400, 280], dims="z", attrs={"units": "K"}), nz,
        horizontal_boundary_type=hb_type, nb=nb, horizontal_boundary_kwargs=hb_kwargs,
        backend="numpy", **kw,
    )


def gen_isentropic_dry(name, nx, ny, nz, scheme, flux, nb, nr, nsteps, dt_s, damp_depth=4,
                       damp_every=True, topo_time=60.0, moist=False, hb_type="relaxed"):
    """Dry isentropic dycore driven through the reference's OWN classes: real Domain /
    Relaxed / RK3WSSI|ForwardEulerSI / IsentropicDiagnostics / Rayleigh / HorizontalVelocity
    and the real ``IsentropicDynamicalCore.stage_array_call_dry``, chained as
    framework/dycore.py:L455-L458 does; after each step the real
    ``get_diagnostic_variables`` refreshes p, exn, mtg, h."""
    d = _make_domain(nx, ny, nz, hb_type, nb, {"nr": nr} if hb_type == "relaxed" else {}, topo_time=topo_time)
    g = d.numerical_grid
    st = refload.load("tasmania.isentropic.state")
    # hb_type == "periodic": (nx, ny) is the PHYSICAL grid, everything below lives on the numerical
    # one, nb ghost points a side (periodic.py:L44-L50); the fixture stores the numerical sizes
    nx, ny = g.nx, g.ny
    shape = (nx + 1, ny + 1, nz + 1)
    state = st.get_isentropic_state_from_brunt_vaisala_frequency(
        g, datetime(2000, 1, 1), da(22.5, "m s^-1"), da(0.0, "m s^-1"), da(0.015, "s^-1"),
        moist=False, backend="numpy", storage_shape=shape,
    )
    if moist:
        # seeded, smooth water species (the moist state builder needs the meteo utilities; the
        # dycore only sees arrays): qv decaying with height, patches of cloud and rain water
        rng = np.random.default_rng(20261020)
        ii, jj, kk = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
        bump = np.exp(-(((ii - 0.45 * nx) / (0.2 * nx)) ** 2) - ((jj - 0.5 * ny) / (0.25 * ny)) ** 2)
        qv = 8e-3 * (kk + 1.0) / (nz + 1.0) * (1.0 + 0.3 * bump) * rng.uniform(0.95, 1.05, size=shape)
        qc = 1e-3 * bump * (kk > nz // 2) * rng.uniform(0.5, 1.0, size=shape)
        qr = 5e-4 * bump * (kk > nz // 3) * rng.uniform(0.0, 1.0, size=shape)
        for key, arr in ((MFWV, qv), (MFCW, qc), (MFPW, qr)):
            arr[nx:, :, :] = 0.0
            arr[:, ny:, :] = 0.0
            arr[:, :, nz:] = 0.0
            state[key] = DataArray(arr, attrs={"units": "g g^-1"})
    hb = d.horizontal_boundary
    hb.reference_state = state

    dyc = refload.load("tasmania.isentropic.dynamics.dycore")
    refload.load("tasmania.isentropic.dynamics.subclasses.prognostics.utils")
    refload.load("tasmania.isentropic.dynamics.subclasses.prognostics.rk3ws_si")
    refload.load("tasmania.isentropic.dynamics.subclasses.prognostics.forward_euler_si")
    for sch in ("upwind", "centered", "third_order_upwind", "fifth_order_upwind"):
        mod = refload.load(
            "tasmania.isentropic.dynamics.subclasses.minimal_horizontal_fluxes." + sch
        )
        if sch == "centered":
            full = refload.load(
                "tasmania.isentropic.dynamics.subclasses.horizontal_fluxes.centered"
            )
            mod.get_centered_flux_x = full.Centered.get_centered_flux_x_numpy
            mod.get_centered_flux_y = full.Centered.get_centered_flux_y_numpy
            # reference quirk: minimal Centered leaves the class attribute ``externals`` at
            # None, so RK3WSSI._stencils_initialize (rk3ws_si.py:L249) raises; give it the
            # empty dict every other scheme effectively has for the numpy backend
            if mod.Centered.externals is None:
                mod.Centered.externals = {}
    refload.load("tasmania.dwarfs.subclasses.vertical_dampers.rayleigh")
    prog = refload.load("tasmania.isentropic.dynamics.prognostic")
    vd = refload.load("tasmania.dwarfs.vertical_damping")
    dd = refload.load("tasmania.dwarfs.diagnostics")
    idiag = refload.load("tasmania.isentropic.dynamics.diagnostics")
    opts = refload.load("tasmania.framework.options")

    pt = float(state[P].data[0, 0, 0])
    bo, so = opts.BackendOptions, opts.StorageOptions
    P_ = prog.IsentropicPrognostic.factory(
        scheme, flux, d, moist, backend="numpy", backend_options=bo(), storage_shape=shape,
        storage_options=so(), pt=da(pt, "Pa"), eps=0.5,
    )
    damper = vd.VerticalDamping.factory(
        "rayleigh", g, damp_depth, 5e-4, backend="numpy", backend_options=bo(),
        storage_shape=shape, storage_options=so(),
    )
    vel = dd.HorizontalVelocity(g, staggering=True, backend="numpy", backend_options=bo(),
                                storage_options=so())
    diags = idiag.IsentropicDiagnostics(g, backend="numpy", backend_options=bo(),
                                        storage_shape=shape, storage_options=so())
    outnames = (S, SU, U, SV, V) + ((MFWV, MFCW, MFPW) if moist else ())
    wc = dd.WaterConstituent(g, clipping=True, backend="numpy", backend_options=bo(),
                             storage_options=so()) if moist else None
    fake = types.SimpleNamespace(
        _water_constituent=wc,
        **({f"_{q}_{t}": np.zeros(shape) for q in ("sqv", "sqc", "sqr") for t in ("now", "int", "new")}
           if moist else {}),
        horizontal_boundary=hb,
        output_properties={k: {"units": state[k].attrs["units"]} for k in outnames},
        _damp=True, _damp_at_every_stage=damp_every, stages=P_.stages, _prognostic=P_,
        _damper=damper, _velocity_components=vel,
        _s_ref=np.zeros(shape), _su_ref=np.zeros(shape), _sv_ref=np.zeros(shape),
        _s_now=None, _su_now=None, _sv_now=None,
    )
    innames = (S, MTG, SU, U, SV, V) + ((MFWV, MFCW, MFPW) if moist else ())
    stage_call = (dyc.IsentropicDynamicalCore.stage_array_call_moist if moist
                  else dyc.IsentropicDynamicalCore.stage_array_call_dry)
    cur = {k: state[k].data.copy() for k in innames}
    cur["time"] = state["time"]
    extra = {k: state[k].data.copy() for k in (P, EXN, H)}
    init = {f"init_{k}": v.copy() for k, v in {**cur, **extra}.items() if k != "time"}
    stage_outs = [{k: np.zeros(shape) for k in outnames} for _ in range(P_.stages)]
    dt = timedelta(seconds=dt_s)
    first_stage = None
    for step in range(nsteps):
        g.update_topography((step + 1) * dt)
        st_in = cur
        for stage in range(P_.stages):
            stage_call(fake, stage, st_in, {}, dt, stage_outs[stage])
            st_in = dict(stage_outs[stage])
            st_in.setdefault(MTG, cur[MTG])
            if step == 0 and stage == 0:
                first_stage = {f"stage0_{k}": stage_outs[0][k].copy() for k in outnames}
        new = {k: stage_outs[-1][k].copy() for k in outnames}
        new["time"] = cur["time"] + dt
        mtg = cur[MTG].copy()
        diags.get_diagnostic_variables(new[S], pt, extra[P], extra[EXN], mtg, extra[H])
        new[MTG] = mtg
        cur = new
    final = {f"final_{k}": v for k, v in {**cur, **extra}.items() if k != "time"}
    save(
        name,
        dims=np.array([nx, ny, nz, nb, nr, nsteps, damp_depth, int(damp_every)]),
        params=np.array([g.dx.to_units("m").values.item(), g.dy.to_units("m").values.item(),
                         g.dz.to_units("K").values.item(), pt, dt_s, 0.5, 5e-4, topo_time]),
        z=g.z.values, z_hl=g.z_on_interface_levels.values,
        x=g.x.to_units("m").values, y=g.y.to_units("m").values,
        topo_steady=g.topography.steady_profile.values,
        gamma=np.array(hb._gamma) if hb_type == "relaxed" else np.zeros(1), rmat=np.array(damper._rmat),
        scheme=np.array([scheme, flux]), boundary=np.array([hb_type]),
        **init, **first_stage, **final,
    )


CASES = {
    "stencils": gen_stencils,
    "kessler": gen_kessler,
    "isentropic_physics": gen_vertical_advection,
    "isen_dry_rk3_5th": lambda: gen_isentropic_dry(
        "isen_dry_rk3_5th", 25, 21, 8, "rk3ws_si", "fifth_order_upwind", 3, 6, 6, 5.0),
    "isen_dry_rk3_3rd": lambda: gen_isentropic_dry(
        "isen_dry_rk3_3rd", 19, 23, 6, "rk3ws_si", "third_order_upwind", 2, 5, 4, 5.0),
    "isen_dry_rk3_cen": lambda: gen_isentropic_dry(
        "isen_dry_rk3_cen", 17, 15, 5, "rk3ws_si", "centered", 1, 4, 3, 4.0, damp_every=False),
    "isen_dry_fe_upw": lambda: gen_isentropic_dry(
        "isen_dry_fe_upw", 16, 18, 5, "forward_euler_si", "upwind", 1, 3, 4, 3.0),
    "isen_moist_rk3_5th": lambda: gen_isentropic_dry(
        "isen_moist_rk3_5th", 23, 19, 8, "rk3ws_si", "fifth_order_upwind", 3, 6, 4, 5.0, moist=True),
    "isen_dry_rk3_5th_periodic": lambda: gen_isentropic_dry(
        "isen_dry_rk3_5th_periodic", 19, 15, 7, "rk3ws_si", "fifth_order_upwind", 3, 0, 5, 5.0,
        damp_depth=3, topo_time=20.0, hb_type="periodic"),
    "isen_dry_rk3_3rd_periodic": lambda: gen_isentropic_dry(
        "isen_dry_rk3_3rd_periodic", 14, 17, 5, "rk3ws_si", "third_order_upwind", 2, 0, 4, 5.0,
        damp_every=False, topo_time=15.0, hb_type="periodic"),
    "isen_dry_fe_cen_periodic": lambda: gen_isentropic_dry(
        "isen_dry_fe_cen_periodic", 13, 16, 5, "forward_euler_si", "centered", 1, 0, 4, 3.0,
        damp_depth=2, topo_time=9.0, hb_type="periodic"),
    "isen_moist_rk3_5th_periodic": lambda: gen_isentropic_dry(
        "isen_moist_rk3_5th_periodic", 17, 16, 6, "rk3ws_si", "fifth_order_upwind", 3, 0, 4, 5.0,
        damp_depth=3, topo_time=20.0, moist=True, hb_type="periodic"),
    "isen_moist_fe_3rd": lambda: gen_isentropic_dry(
        "isen_moist_fe_3rd", 17, 21, 6, "forward_euler_si", "third_order_upwind", 2, 5, 3, 4.0,
        moist=True),
}


CASES["stencils_1d"] = gen_stencils_1d


if __name__ == "__main__":
    todo = sys.argv[1:] or list(CASES)
    for case in todo:
        CASES[case]()
