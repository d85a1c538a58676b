# -*- coding: utf-8 -*-
"""The recording C-ABI stub (tests/abi_stub.py) with the ORACLE behind every entry point the
isentropic models use (test infrastructure, CPU only).

Each ``tb200_*`` call that reaches the stub is decoded exactly as the C side decodes it -- field
pointers with their shapes and strides (sliced and zero-stride views included), scalars, flag
words, origin / domain boxes -- and carried out on the host buffers by the numpy oracle function
that restates the corresponding reference stencil.  The whole host side of the product (mirrors,
coupling layer, model assembly, argument marshalling, buffer rotation, CUDA-graph bookkeeping) can
thereby be run *numerically* in the GPU-less container and compared with the oracle's own models;
what this cannot check is the CUDA code, which the ``-m gpu`` tests hold to the same oracle.
"""
import numpy as np

from oracle import boundary as ob
from oracle import burgers as obu
from oracle import dwarfs
from oracle import isentropic as oi
from oracle import isentropic_physics as op
from oracle import microphysics as om
from tasmania_b200 import lib
from tests.abi_stub import AbiStub, _as_numpy, _box

FLUX_NAMES = {v: k for k, v in lib.FLUX_SCHEMES.items()}
OPS = {v: k for k, v in lib.ELEMENTWISE_OPS.items()}
KF = lib.KESSLER_FLAGS


def arr(field_p):
    """numpy view of a ``tb200_field*`` (None for NULL)."""
    return _as_numpy(field_p.contents) if field_p else None


def box(o, d):
    return tuple(int(o[n]) for n in range(3)), tuple(int(d[n]) for n in range(3))


def three(pp):
    """An array of three field pointers (qv, qc, qr) -> list of views, or None for NULL."""
    return [arr(pp[n]) for n in range(3)] if pp else None


def constants(c):
    return {"pref": float(c[0]), "rd": float(c[1]), "g": float(c[2]), "cp": float(c[3])}


class OracleStub(AbiStub):
    # ---- K12
    def _do_tb200_elementwise(self, code, out, a, b, c, f, o, d, stream):
        bx = _box(o, d)
        va, vb, vc = (None if x is None else x[bx] for x in (arr(a), arr(b), arr(c)))
        res = {
            "copy": lambda: va, "copychange": lambda: -va, "abs": lambda: np.abs(va),
            "add": lambda: va + vb, "addsub": lambda: va + vb - vc,
            "clip": lambda: np.where(va > 0, va, 0), "fma": lambda: va + f * vb, "mul": lambda: va * vb,
            "scale": lambda: f * va, "sub": lambda: va - vb,
            "sts_rk2_0": lambda: 0.5 * (va + vb + f * vc),
            "sts_rk3ws_0": lambda: (2.0 * va + vb + f * vc) / 3.0,
            "iaddsub": lambda: va + (vb - vc), "iscale": lambda: va * f,
        }[OPS[code]]()
        arr(out)[bx] = res

    # ---- K5
    def _do_tb200_relax(self, gamma, phi, ref, out, o, d, stream):
        origin, domain = box(o, d)
        if phi:
            ob.relax(arr(gamma), arr(phi), arr(ref), arr(out), origin, domain)
        else:
            ob.irelax(arr(gamma), arr(ref), arr(out), origin, domain)

    def _do_tb200_set_outermost_layers(self, field, ref, axis, mi, mj, stream):
        f, r = arr(field), arr(ref)
        if axis == 0:  # relaxed.py:L161-L175
            f[0, :mj] = r[0, :mj]
            f[mi - 1, :mj] = r[mi - 1, :mj]
        else:  # L177-L191
            f[:mi, 0] = r[:mi, 0]
            f[:mi, mj - 1] = r[:mi, mj - 1]

    def _do_tb200_periodic_enforce(self, field, nx, ny, nb, mx, my, stream):
        stag = {(True, True): "at_uv_locations", (True, False): "at_u_locations",
                (False, True): "at_v_locations", (False, False): ""}[(mx > nx, my > ny)]
        ob.Periodic(nx, ny, 1, nb).enforce_field(arr(field), stag)

    # ---- K10
    def _do_tb200_burgers_forward_euler(self, order, u, v, u_tmp, v_tmp, out_u, out_v, u_tnd, v_tnd, dt, dx, dy,
                                        o, d, stream):
        origin, domain = box(o, d)
        obu.forward_euler(order, arr(u), arr(v), arr(u_tmp), arr(v_tmp), arr(out_u), arr(out_v), dt=dt, dx=dx,
                          dy=dy, origin=origin, domain=domain, u_tnd=arr(u_tnd), v_tnd=arr(v_tnd))

    # ---- K6, K4, K7
    def _do_tb200_damping(self, now, new, ref, rmat, out, dt, o, d, stream):
        dwarfs.damping(arr(now), arr(new), arr(ref), arr(rmat), arr(out), dt, *box(o, d))

    def _do_tb200_velocity(self, axis, dens, mom, out, staggering, o, d, stream):
        fn = dwarfs.velocity_x if axis == 0 else dwarfs.velocity_y
        fn(arr(dens), arr(mom), arr(out), *box(o, d), staggering=bool(staggering))

    def _do_tb200_density(self, dens, q, dq, clipping, o, d, stream):
        dwarfs.density(arr(dens), arr(q), arr(dq), *box(o, d), clipping=bool(clipping))

    def _do_tb200_mass_fraction(self, dens, dq, q, clipping, o, d, stream):
        dwarfs.mass_fraction(arr(dens), arr(dq), arr(q), *box(o, d), clipping=bool(clipping))

    def _do_tb200_thomas(self, a, b, c, d, out, o, dm, stream):
        op.thomas(arr(a), arr(b), arr(c), arr(d), arr(out), *box(o, dm))

    # ---- K8, K9
    def _do_tb200_diffusion(self, order, phi, gamma, out, dx, dy, ow, o, d, stream):
        dwarfs.diffusion(order, arr(phi), arr(gamma), arr(out), dx, dy, bool(ow), *box(o, d))

    def _do_tb200_hyperdiffusion(self, phi, out, alpha, o, d, stream):
        dwarfs.hyperdiffusion(arr(phi), arr(out), alpha, *box(o, d))

    def _do_tb200_diffusion_1d(self, order, axis, phi, gamma, out, h, ow, o, d, stream):
        dwarfs.diffusion_1d(order, axis, arr(phi), arr(gamma), arr(out), h, bool(ow), *box(o, d))

    def _do_tb200_smoothing_1d(self, order, axis, phi, gamma, out, rim_copy, o, d, stream):
        origin, domain = box(o, d)
        vin, vout = arr(phi), arr(out)
        dwarfs.smoothing_1d(order, axis, vin, arr(gamma), vout, origin, domain)
        if rim_copy:  # the two copies of a 1-D smoother's __call__, first_order.py:L187-L204
            (i0, j0, k0), (di, dj, dk) = origin, domain
            ni, nj, k = 2 * i0 + di, 2 * j0 + dj, slice(k0, k0 + dk)
            rim = np.ones((ni, nj), dtype=bool)
            rim[i0:i0 + di, j0:j0 + dj] = False
            vout[:ni, :nj, k] = np.where(rim[:, :, None], vin[:ni, :nj, k], vout[:ni, :nj, k])

    def _do_tb200_smoothing(self, order, phi, gamma, out, rim_copy, o, d, stream):
        origin, domain = box(o, d)
        vin, vout = arr(phi), arr(out)
        dwarfs.smoothing(order, vin, arr(gamma), vout, origin, domain)
        if rim_copy:  # the four copies of HorizontalSmoothing.__call__, first_order.py:L77-L110
            (i0, j0, k0), (di, dj, dk) = origin, domain
            ni, nj, k = 2 * i0 + di, 2 * j0 + dj, slice(k0, k0 + dk)
            rim = np.ones((ni, nj), dtype=bool)
            rim[i0:i0 + di, j0:j0 + dj] = False
            vout[:ni, :nj, k] = np.where(rim[:, :, None], vin[:ni, :nj, k], vout[:ni, :nj, k])

    # ---- K1, K2, K3
    def _do_tb200_step_forward_euler(self, flux, s_now, s_int, s_new, u, v, s_tnd, sq_now, sq_int, sq_new,
                                     q_tnd, dt, dx, dy, o, d, stream):
        origin, domain = box(o, d)
        kw = {}
        if sq_now:
            kw = dict(moist=True, sq_now=three(sq_now), sq_int=three(sq_int), sq_new=three(sq_new),
                      q_tnd=three(q_tnd) or (None, None, None))
        oi.step_forward_euler(FLUX_NAMES[flux], arr(s_now), arr(s_int), arr(s_new), arr(u), arr(v), dt=dt,
                              dx=dx, dy=dy, origin=origin, domain=domain, s_tnd=arr(s_tnd), **kw)

    def _do_tb200_step_forward_euler_momentum(self, flux, s_now, s_new, u, v, su_now, su_int, su_new, sv_now,
                                              sv_int, sv_new, mtg_now, mtg_new, su_tnd, sv_tnd, dt, dx, dy,
                                              eps, o, d, stream):
        origin, domain = box(o, d)
        oi.step_forward_euler_momentum(
            FLUX_NAMES[flux], arr(s_now), arr(s_new), arr(u), arr(v), arr(su_now), arr(su_int), arr(su_new),
            arr(sv_now), arr(sv_int), arr(sv_new), arr(mtg_now), arr(mtg_new), dt=dt, dx=dx, dy=dy, eps=eps,
            origin=origin, domain=domain, su_tnd=arr(su_tnd), sv_tnd=arr(sv_tnd))

    def _do_tb200_montgomery(self, hs, s, mtg, dz, pt, theta_s, c, o, d, stream):
        origin, domain = box(o, d)
        oi.montgomery(arr(hs), arr(s), arr(mtg), dz=dz, pt=pt, theta_s=theta_s, origin=origin, domain=domain,
                      constants=constants(c))

    def _do_tb200_diagnostic_variables(self, theta, hs, s, p, exn, mtg, h, dz, pt, c, o, d, stream):
        origin, domain = box(o, d)
        oi.diagnostic_variables(arr(theta), arr(hs), arr(s), arr(p), arr(exn), arr(mtg), arr(h), dz=dz, pt=pt,
                                origin=origin, domain=domain, constants=constants(c))

    def _do_tb200_height(self, theta, hs, s, h, dz, pt, c, o, d, stream):
        origin, domain = box(o, d)
        oi.height(arr(theta), arr(hs), arr(s), arr(h), dz=dz, pt=pt, origin=origin, domain=domain,
                  constants=constants(c))

    def _do_tb200_density_and_temperature(self, theta, s, exn, h, rho, t, cp, o, d, stream):
        origin, domain = box(o, d)
        oi.density_and_temperature(arr(theta), arr(s), arr(exn), arr(h), arr(rho), arr(t), origin=origin,
                                   domain=domain, constants=dict(oi.CONSTANTS, cp=cp))

    # ---- the fused dry RK stage (csrc/isentropic_fused.cu): what dycore.py:L641-L721 over
    # rk3ws_si.py:L105-L234 computes for one stage, from the same arguments the kernels get
    def _do_tb200_isentropic_stage_moist(self, cfg, s_now, su_now, sv_now, mtg_now, s_int, su_int, sv_int, u_int,
                                         v_int, s_new, su_new, sv_new, u_new, v_new, s_ref, su_ref, sv_ref,
                                         u_ref, v_ref, gamma2d, rmat, topo2d, scr0, scr1, scr2, q_now, q_int,
                                         q_new, q_ref, stream):
        """dycore.py:L723-L843: the dry stage plus, per water constituent, density (now, int) -> K1 ->
        mass fraction over the stage's s after its first relaxation -> lateral relaxation."""
        c = cfg.contents
        nx, ny, nz, nb = c.nx, c.ny, c.nz, c.nb

        def tracers(s_pre, u_i, v_i, gamma):
            flux = FLUX_NAMES[c.flux_scheme]
            origin, domain = (nb, nb, 0), (nx - 2 * nb, ny - 2 * nb, nz)
            full = ((0, 0, 0), (nx, ny, nz))
            for t in range(3):
                qn, qi, qo, qr = (arr(x[t]) for x in (q_now, q_int, q_new, q_ref))
                sq_now, sq_int, sq_new = np.zeros_like(qn), np.zeros_like(qn), np.zeros_like(qn)
                dwarfs.density(arr(s_now), qn, sq_now, *full)
                dwarfs.density(arr(s_int), qi, sq_int, *full)
                with np.errstate(all="ignore"):
                    oi.step_forward_euler(flux, arr(s_now), arr(s_int), np.zeros_like(qn), u_i, v_i, dt=c.dt,
                                          dx=c.dx, dy=c.dy, origin=origin, domain=domain, moist=True,
                                          sq_now=[sq_now], sq_int=[sq_int], sq_new=[sq_new], q_tnd=(None,))
                    new = qo.copy()
                    dwarfs.mass_fraction(s_pre, sq_new, new, *full)
                box = (slice(nb, nx - nb), slice(nb, ny - nb), slice(0, nz))
                qo[box] = new[box]  # outside: the reference divides stale sq_new; gamma == 1 there
                ob.irelax(gamma, qr, qo, (0, 0, 0), (nx, ny, nz))

        self._do_tb200_isentropic_stage_dry(cfg, s_now, su_now, sv_now, mtg_now, s_int, su_int, sv_int, u_int,
                                            v_int, s_new, su_new, sv_new, u_new, v_new, s_ref, su_ref, sv_ref,
                                            u_ref, v_ref, gamma2d, rmat, topo2d, scr0, scr1, scr2, stream,
                                            after_s_step=tracers)

    def _do_tb200_isentropic_stage_dry(self, cfg, s_now, su_now, sv_now, mtg_now, s_int, su_int, sv_int, u_int,
                                       v_int, s_new, su_new, sv_new, u_new, v_new, s_ref, su_ref, sv_ref,
                                       u_ref, v_ref, gamma2d, rmat, topo2d, scr0, scr1, scr2, stream,
                                       after_s_step=None):
        c = cfg.contents
        assert c.part == 0, "the overlap parts of a decomposed run are not emulated"
        nx, ny, nz, nb = c.nx, c.ny, c.nz, c.nb
        flux = FLUX_NAMES[c.flux_scheme]
        out = {n: arr(x) for n, x in (("s", s_new), ("su", su_new), ("sv", sv_new), ("u", u_new), ("v", v_new))}
        shape = out["s"].shape
        gamma = np.broadcast_to(arr(gamma2d)[:, :, :1], shape)
        origin, domain = (nb, nb, 0), (nx - 2 * nb, ny - 2 * nb, nz)
        u_i, v_i = arr(u_int), arr(v_int)
        if c.derive_uv_in:  # the kernels re-diagnose the advecting velocities; u_int / v_int may be stale
            u_i, v_i = np.full(shape, np.nan), np.full(shape, np.nan)
            dwarfs.get_velocity_components(nx, ny, nz, arr(s_int), arr(su_int), arr(sv_int), u_i, v_i)
        assert arr(scr2) is not None  # may BE s_new (in-place update with skip_uv_out)
        assert c.skip_uv_out or arr(scr2).ctypes.data != out["s"].ctypes.data
        s_tnd, su_tnd, sv_tnd = (arr(x) if x else None for x in (c.s_tnd, c.su_tnd, c.sv_tnd))
        assert (s_tnd is None) == (su_tnd is None) == (sv_tnd is None)  # all three or none
        oi.step_forward_euler(flux, arr(s_now), arr(s_int), out["s"], u_i, v_i, dt=c.dt,
                              dx=c.dx, dy=c.dy, origin=origin, domain=domain, s_tnd=s_tnd)
        ob.irelax(gamma, arr(s_ref), out["s"], (0, 0, 0), (nx, ny, nz))
        if c.periodic:  # the stage wraps s itself (hb.enforce_field(s_new), rk3ws_si.py:L184-L189)
            assert not gamma.any() and not c.damp and c.skip_uv_out
            ob.Periodic(nx - 2 * nb, ny - 2 * nb, nz, nb).enforce_field(out["s"][:nx, :ny, :nz])
        if after_s_step is not None:
            after_s_step(out["s"], u_i, v_i, gamma)
        hs = np.zeros(shape)
        hs[:, :, nz] = arr(topo2d)[: shape[0], : shape[1], 0]
        mtg_new = np.zeros(shape)
        oi.montgomery(hs, out["s"], mtg_new, dz=c.dz, pt=c.pt, theta_s=c.theta_s, origin=(0, 0, 0),
                      domain=(nx, ny, nz + 1), constants=constants(c.constants))
        oi.step_forward_euler_momentum(
            flux, arr(s_now), out["s"], u_i, v_i, arr(su_now), arr(su_int), out["su"], arr(sv_now),
            arr(sv_int), out["sv"], arr(mtg_now), mtg_new, dt=c.dt, dx=c.dx, dy=c.dy, eps=c.eps, origin=origin,
            domain=domain, su_tnd=su_tnd, sv_tnd=sv_tnd)
        for n, ref in (("s", s_ref), ("su", su_ref), ("sv", sv_ref)):  # enforce_raw (u, v are re-diagnosed)
            ob.irelax(gamma, arr(ref), out[n], (0, 0, 0), (nx, ny, nz))
        if c.damp:
            r = np.broadcast_to(arr(rmat)[:1, :1, :], shape)
            for n, now, ref in (("s", s_now, s_ref), ("su", su_now, su_ref), ("sv", sv_now, sv_ref)):
                dwarfs.damping(arr(now), out[n], arr(ref), r, out[n], c.dt_full, (0, 0, 0), shape)
        if c.skip_uv_out:  # not written by the kernels: poison, so that any consumer shows up
            out["u"][: nx + 1, :ny, :nz] = np.nan
            out["v"][:nx, : ny + 1, :nz] = np.nan
            return
        dwarfs.get_velocity_components(nx, ny, nz, out["s"], out["su"], out["sv"], out["u"], out["v"])
        ur, vr = arr(u_ref), arr(v_ref)
        out["u"][0, :ny], out["u"][nx, :ny] = ur[0, :ny], ur[nx, :ny]        # relaxed.py:L161-L175
        out["v"][:nx, 0], out["v"][:nx, ny] = vr[:nx, 0], vr[:nx, ny]        # L177-L191

    def _do_tb200_velocity_components(self, d, du, dv, u, v, u_ref, v_ref, nx, ny, nz, stream):
        """dwarfs/diagnostics.py:L219-L272 + relaxed.py:L161-L191 (csrc/elementwise.cu)."""
        uo, vo = arr(u), arr(v)
        dwarfs.get_velocity_components(nx, ny, nz, arr(d), arr(du), arr(dv), uo, vo)
        if u_ref:
            ur = arr(u_ref)
            uo[0, :ny, :nz], uo[nx, :ny, :nz] = ur[0, :ny, :nz], ur[nx, :ny, :nz]
        if v_ref:
            vr = arr(v_ref)
            vo[:nx, 0, :nz], vo[:nx, ny, :nz] = vr[:nx, 0, :nz], vr[:nx, ny, :nz]

    # ---- K11
    def _do_tb200_kessler(self, rho, p, t, exn, qc, qr, qv, t_qc, t_qr, t_qv, t_th, a, k1, k2, beta, lhvw, flags,
                          o, d, stream):
        origin, domain = box(o, d)
        om.kessler(arr(rho), arr(p), arr(t), arr(exn), arr(qc), arr(qr), arr(qv), arr(t_qc), arr(t_qr),
                   arr(t_qv), arr(t_th), a=a, k1=k1, k2=k2, ow_out_qc_tnd=bool(flags & KF["ow_qc"]),
                   ow_out_qr_tnd=bool(flags & KF["ow_qr"]), ow_out_qv_tnd=bool(flags & KF["ow_qv"]),
                   ow_out_theta_tnd=bool(flags & KF["ow_theta"]), origin=origin, domain=domain,
                   air_pressure_on_interface_levels=bool(flags & KF["p_on_interfaces"]),
                   rain_evaporation=bool(flags & KF["rain_evaporation"]), beta=beta, lhvw=lhvw)

    def _do_tb200_saturation_prognostic(self, p, t, exn, qv, qc, t_qv, t_qc, t_th, sr, beta, lhvw, cp, rv, flags,
                                        o, d, stream):
        origin, domain = box(o, d)
        om.saturation_prognostic(
            arr(p), arr(t), arr(exn), arr(qv), arr(qc), arr(t_qv), arr(t_qc), arr(t_th), sr=sr, origin=origin,
            domain=domain, ow_tnd_qv=bool(flags & KF["ow_qv"]), ow_tnd_qc=bool(flags & KF["ow_qc"]),
            ow_tnd_theta=bool(flags & KF["ow_theta"]),
            air_pressure_on_interface_levels=bool(flags & KF["p_on_interfaces"]), beta=beta, lhvw=lhvw,
            cp=cp, rv=rv)

    def _do_tb200_fall_velocity(self, rho, rho_s, qr, vt, o, d, stream):
        origin, domain = box(o, d)
        om.fall_velocity(arr(rho), arr(rho_s), arr(qr), arr(vt), origin=origin, domain=domain)

    def _do_tb200_sedimentation(self, order, rho, h, qr, vt, tnd, ow, o, d, stream):
        origin, domain = box(o, d)
        om.sedimentation(arr(rho), arr(h), arr(qr), arr(vt), arr(tnd), ow_out_tnd_qr=bool(ow), origin=origin,
                         domain=domain, order=order)

    def _do_tb200_accumulated_precipitation(self, rho, qr, vt, acc, prec, out_acc, dt, rhow, o, d, stream):
        origin, domain = box(o, d)
        om.accumulated_precipitation(arr(rho), arr(qr), arr(vt), arr(acc), arr(prec), arr(out_acc), dt=dt,
                                     origin=origin, domain=domain, rhow=rhow)

    # ---- isentropic physics
    def _do_tb200_coriolis(self, su, sv, t_su, t_sv, f, ow_su, ow_sv, o, d, stream):
        origin, domain = box(o, d)
        op.coriolis(arr(su), arr(sv), arr(t_su), arr(t_sv), f=f, ow_tnd_su=bool(ow_su), ow_tnd_sv=bool(ow_sv),
                    origin=origin, domain=domain)

    def _do_tb200_smagorinsky(self, s, a, b, t_a, t_b, dx, dy, cs, ow_a, ow_b, o, d, stream):
        origin, domain = box(o, d)
        with np.errstate(divide="ignore", invalid="ignore"):
            op.smagorinsky(arr(a), arr(b), arr(t_a), arr(t_b), dx=dx, dy=dy, cs=cs, ow_out_u_tnd=bool(ow_a),
                           ow_out_v_tnd=bool(ow_b), origin=origin, domain=domain, in_s=arr(s))

    def _do_tb200_coriolis_step(self, su, sv, b_su, b_sv, o_su, o_sv, f, factor, o, d, full, stream):
        origin, domain = box(o, d)
        shape = tuple(int(full[n]) for n in range(3))
        t_su, t_sv = np.zeros(arr(o_su).shape), np.zeros(arr(o_sv).shape)
        op.coriolis(arr(su), arr(sv), t_su, t_sv, f=f, ow_tnd_su=True, ow_tnd_sv=True, origin=origin, domain=domain)
        fb = tuple(slice(0, n) for n in shape)
        arr(o_su)[fb] = arr(b_su)[fb] + factor * t_su[fb]
        arr(o_sv)[fb] = arr(b_sv)[fb] + factor * t_sv[fb]

    def _do_tb200_smagorinsky_step(self, s, a, b, b_a, b_b, o_a, o_b, dx, dy, cs, factor, o, d, full, stream):
        origin, domain = box(o, d)
        shape = tuple(int(full[n]) for n in range(3))
        t_a, t_b = np.zeros(arr(o_a).shape), np.zeros(arr(o_b).shape)
        with np.errstate(divide="ignore", invalid="ignore"):
            op.smagorinsky(arr(a), arr(b), t_a, t_b, dx=dx, dy=dy, cs=cs, ow_out_u_tnd=True, ow_out_v_tnd=True,
                           origin=origin, domain=domain, in_s=arr(s))
        fb = tuple(slice(0, n) for n in shape)
        arr(o_a)[fb] = arr(b_a)[fb] + factor * t_a[fb]
        arr(o_b)[fb] = arr(b_b)[fb] + factor * t_b[fb]

    def _vadv(self, flux, staggered, w, ins, outs, dz, ows, o, d):
        origin, domain = box(o, d)
        moist = len(ins) == 6
        kw = {}
        if moist:
            kw = dict(in_qv=ins[3], in_qc=ins[4], in_qr=ins[5], out_qv=outs[3], out_qc=outs[4], out_qr=outs[5],
                      ow_out_qv=ows[3], ow_out_qc=ows[4], ow_out_qr=ows[5])
        op.vertical_advection(FLUX_NAMES[flux], bool(staggered), w, ins[0], ins[1], ins[2], outs[0], outs[1],
                              outs[2], dz=dz, ow_out_s=ows[0], ow_out_su=ows[1], ow_out_sv=ows[2],
                              origin=origin, domain=domain, **kw)

    def _do_tb200_vertical_advection(self, flux, staggered, w, s, su, sv, o_s, o_su, o_sv, qv, qc, qr, o_qv,
                                     o_qc, o_qr, dz, flags, o, d, stream):
        ins = [arr(x) for x in ((s, su, sv, qv, qc, qr) if qv else (s, su, sv))]
        outs = [arr(x) for x in ((o_s, o_su, o_sv, o_qv, o_qc, o_qr) if qv else (o_s, o_su, o_sv))]
        self._vadv(flux, staggered, arr(w), ins, outs, dz, [bool((flags >> n) & 1) for n in range(6)], o, d)

    def _do_tb200_vertical_advection_step(self, flux, staggered, w, n, ins, bases, outs, dz, factor, o, d, stream):
        vin = [arr(ins[m]) for m in range(n)]
        vbase = [arr(bases[m]) for m in range(n)]
        vout = [arr(outs[m]) for m in range(n)]
        tnd = [np.zeros_like(x) for x in vout]
        self._vadv(flux, staggered, arr(w), vin, tnd, dz, [True] * 6, o, d)
        for m in range(n):  # the stage update over the whole output storage
            vout[m][...] = vbase[m][tuple(slice(0, s) for s in vout[m].shape)] + factor * tnd[m]
