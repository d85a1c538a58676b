# -*- coding: utf-8 -*-
"""Oracle (test infrastructure): the isentropic physics components of SURVEY.md 8f -- vertical
advection (row 1), Coriolis forcing and Smagorinsky turbulence (row 3), implicit vertical
advection with the Thomas algorithm (row 4).

Follows src/tasmania/isentropic/physics/vertical_advection.py:L271-L386 (numpy definition of
``IsentropicVerticalAdvection._stencil``) and the flux formulas of
src/tasmania/isentropic/dynamics/subclasses/minimal_vertical_fluxes/{upwind.py:L31-L33,
centered.py:L28-L30, third_order_upwind.py:L31-L38, fifth_order_upwind.py:L31-L42}.
Pinned bit for bit on tests/golden/isentropic_physics.npz (the reference's own code run in place).
"""
import numpy as np

EXTENT = {"upwind": 1, "centered": 1, "third_order_upwind": 2, "fifth_order_upwind": 3}


def vertical_flux(scheme, w, phi):
    """Flux through the interfaces e .. n - e of a column with n levels (``phi``: n levels,
    ``w``: n + 1 interfaces, third axis); interface K separates the levels K - 1 and K."""
    n = phi.shape[2]
    e = EXTENT[scheme]
    K = np.arange(e, n - e + 1)

    def lev(off):
        return phi[:, :, K + off]

    wk = w[:, :, K]
    if scheme == "upwind":
        return wk * np.where(wk > 0.0, lev(0), lev(-1))
    if scheme == "centered":
        return wk * 0.5 * (lev(0) + lev(-1))
    if scheme == "third_order_upwind":
        return wk / 12.0 * (7.0 * (lev(-1) + lev(0)) - (lev(-2) + lev(1))) - np.abs(wk) / 12.0 * (
            3.0 * (lev(-1) - lev(0)) - (lev(-2) - lev(1)))
    if scheme == "fifth_order_upwind":
        return wk / 60.0 * (
            37.0 * (lev(-1) + lev(0)) - 8.0 * (lev(-2) + lev(1)) + (lev(-3) + lev(2))
        ) - np.abs(wk) / 60.0 * (
            10.0 * (lev(-1) - lev(0)) - 5.0 * (lev(-2) - lev(1)) + (lev(-3) - lev(2)))
    raise ValueError(scheme)


def _set_output(lhs, rhs, overwrite):  # generics.py:L38-L40: on the WHOLE storage
    lhs[...] = rhs if overwrite else lhs + rhs


def vertical_advection(scheme, staggered, in_w, in_s, in_su, in_sv, out_s, out_su, out_sv, *,
                       dz, ow_out_s, ow_out_su, ow_out_sv, origin, domain, in_qv=None, in_qc=None,
                       in_qr=None, out_qv=None, out_qc=None, out_qr=None, ow_out_qv=True,
                       ow_out_qc=True, ow_out_qr=True):
    i = slice(origin[0], origin[0] + domain[0])
    j = slice(origin[1], origin[1] + domain[1])
    kb, ke = origin[2], origin[2] + domain[2]
    e = EXTENT[scheme]
    if staggered:  # L305-L306
        w = in_w
    else:  # L307-L320: averaged onto the inner interfaces, zero on the outermost two
        w = np.zeros(tuple(max(a, b) for a, b in zip(in_w.shape, (0, 0, ke + 1))), dtype=in_w.dtype)
        w[i, j, kb + 1:ke] = 0.5 * (in_w[i, j, kb + 1:ke] + in_w[i, j, kb:ke - 1])
    wcol = w[i, j, kb:ke + 1]

    def tendency(phi, denom):
        f = vertical_flux(scheme, wcol, phi)
        tmp = np.zeros_like(in_s)
        tmp[i, j, kb + e:ke - e] = (f[:, :, 1:] - f[:, :, :-1]) / denom
        return tmp

    for src, out, ow in ((in_s, out_s, ow_out_s), (in_su, out_su, ow_out_su), (in_sv, out_sv, ow_out_sv)):
        _set_output(out, tendency(src[i, j, kb:ke], dz), ow)  # L341-L353
    if in_qv is not None:  # L322-L329, L355-L385
        s_in = in_s[i, j, kb + e:ke - e]
        for q, out, ow in ((in_qv, out_qv, ow_out_qv), (in_qc, out_qc, ow_out_qc), (in_qr, out_qr, ow_out_qr)):
            sq = in_s[i, j, kb:ke] * q[i, j, kb:ke]
            _set_output(out, tendency(sq, s_in * dz), ow)


def coriolis(in_su, in_sv, tnd_su, tnd_sv, *, f, ow_tnd_su, ow_tnd_sv, origin, domain):
    """src/tasmania/isentropic/physics/coriolis.py:L166-L186 (set_output on the box)."""
    box = tuple(slice(o, o + d) for o, d in zip(origin, domain))
    _set_output(tnd_su[box], f * in_sv[box], ow_tnd_su)
    _set_output(tnd_sv[box], -f * in_su[box], ow_tnd_sv)


def _smagorinsky_core(u, v, dx, dy, cs, ib, ie, jb, je, k):
    """src/tasmania/physics/turbulence.py:L211-L229: strain rates on the box + 1 from centred
    differences, eddy viscosity, divergence of the stresses on the box."""
    def dxc(f, i0, i1, j0, j1):
        return (f[i0 + 1:i1 + 1, j0:j1, k] - f[i0 - 1:i1 - 1, j0:j1, k]) / (2.0 * dx)

    def dyc(f, i0, i1, j0, j1):
        return (f[i0:i1, j0 + 1:j1 + 1, k] - f[i0:i1, j0 - 1:j1 - 1, k]) / (2.0 * dy)

    e = (ib - 1, ie + 1, jb - 1, je + 1)
    s00 = dxc(u, *e)
    s01 = 0.5 * (dyc(u, *e) + dxc(v, *e))
    s11 = dyc(v, *e)
    nu = cs**2 * dx * dy * np.sqrt(2.0 * (s00 * s00 + 2.0 * (s01 * s01) + s11 * s11))
    p00, p01, p11 = nu * s00, nu * s01, nu * s11
    u_tnd = 2.0 * ((p00[2:, 1:-1] - p00[:-2, 1:-1]) / (2.0 * dx) + (p01[1:-1, 2:] - p01[1:-1, :-2]) / (2.0 * dy))
    v_tnd = 2.0 * ((p01[2:, 1:-1] - p01[:-2, 1:-1]) / (2.0 * dx) + (p11[1:-1, 2:] - p11[1:-1, :-2]) / (2.0 * dy))
    return u_tnd, v_tnd


def smagorinsky(in_u, in_v, out_u_tnd, out_v_tnd, *, dx, dy, cs, ow_out_u_tnd, ow_out_v_tnd, origin,
                domain, in_s=None):
    """Smagorinsky2d (in_s None) / IsentropicSmagorinsky (in_u, in_v = su, sv; tendencies of them):
    physics/turbulence.py:L165-L187, isentropic/physics/turbulence.py:L99-L125."""
    ib, ie = origin[0], origin[0] + domain[0]
    jb, je = origin[1], origin[1] + domain[1]
    k = slice(origin[2], origin[2] + domain[2])
    if in_s is not None:
        with np.errstate(divide="ignore", invalid="ignore"):
            u, v = in_u / in_s, in_v / in_s
    else:
        u, v = in_u, in_v
    tu, tv = _smagorinsky_core(u, v, dx, dy, cs, ib, ie, jb, je, k)
    if in_s is not None:
        tu, tv = in_s[ib:ie, jb:je, k] * tu, in_s[ib:ie, jb:je, k] * tv
    _set_output(out_u_tnd[ib:ie, jb:je, k], tu, ow_out_u_tnd)
    _set_output(out_v_tnd[ib:ie, jb:je, k], tv, ow_out_v_tnd)


def _setup_tridiagonal(gamma, w, phi):
    """cla.py:L81-L108 on columns (third axis = the nk levels of the box); b is one."""
    a, c, d = np.zeros_like(phi), np.zeros_like(phi), phi.copy()
    a[:, :, 1:-1] = gamma * w[:, :, :-2]
    c[:, :, 1:-1] = -gamma * w[:, :, 2:]
    d[:, :, 1:-1] = phi[:, :, 1:-1] - gamma * (w[:, :, :-2] * phi[:, :, :-2] - w[:, :, 2:] * phi[:, :, 2:])
    return a, c, d


def thomas(a, b, c, d, out, origin, domain):
    """The global ``thomas`` stencil, framework/subclasses/stencil_definitions/cla.py:L33-L62:
    general diagonal b, the reference's zero-pivot rule, on copies of b and d."""
    i, j = (slice(o, o + n) for o, n in zip(origin[:2], domain[:2]))
    k0, k1 = origin[2], origin[2] + domain[2]
    beta, delta = b.copy(), d.copy()
    with np.errstate(divide="ignore", invalid="ignore"):
        for k in range(k0 + 1, k1):
            w = np.where(beta[i, j, k - 1] != 0.0, a[i, j, k] / beta[i, j, k - 1], a[i, j, k])
            beta[i, j, k] -= w * c[i, j, k - 1]
            delta[i, j, k] -= w * delta[i, j, k - 1]
        out[i, j, k1 - 1] = np.where(beta[i, j, k1 - 1] != 0.0, delta[i, j, k1 - 1] / beta[i, j, k1 - 1],
                                     delta[i, j, k1 - 1] / b[i, j, k1 - 1])
        for k in range(k1 - 2, k0 - 1, -1):
            r = delta[i, j, k] - c[i, j, k] * out[i, j, k + 1]
            out[i, j, k] = np.where(beta[i, j, k] != 0.0, r / beta[i, j, k], r / b[i, j, k])


def _thomas(a, c, d):
    """cla.py:L42-L78 with b = 1: forward elimination, backward substitution, level by level."""
    nk = d.shape[2]
    beta, delta = np.ones_like(d), d.copy()
    for k in range(1, nk):
        w = np.where(beta[:, :, k - 1] != 0.0, a[:, :, k] / beta[:, :, k - 1], a[:, :, k])
        beta[:, :, k] -= w * c[:, :, k - 1]
        delta[:, :, k] -= w * delta[:, :, k - 1]
    out = np.empty_like(d)
    out[:, :, -1] = np.where(beta[:, :, -1] != 0.0, delta[:, :, -1] / beta[:, :, -1], delta[:, :, -1] / 1.0)
    for k in range(nk - 2, -1, -1):
        r = delta[:, :, k] - c[:, :, k] * out[:, :, k + 1]
        out[:, :, k] = np.where(beta[:, :, k] != 0.0, r / beta[:, :, k], r / 1.0)
    return out


def implicit_vertical_advection(staggered, in_w, in_s, in_su, in_sv, out_s, out_su, out_sv, *, gamma,
                                origin, domain, in_qv=None, in_qc=None, in_qr=None, out_qv=None,
                                out_qc=None, out_qr=None):
    """implicit_vertical_advection.py:L221-L336."""
    box = tuple(slice(o, o + d) for o, d in zip(origin, domain))
    i, j, _ = box
    kb, ke = origin[2], origin[2] + domain[2]
    w = 0.5 * (in_w[i, j, kb:ke] + in_w[i, j, kb + 1:ke + 1]) if staggered else in_w[box]
    with np.errstate(all="ignore"):
        for src, out in ((in_s, out_s), (in_su, out_su), (in_sv, out_sv)):
            out[box] = _thomas(*_setup_tridiagonal(gamma, w, src[box]))
        if in_qv is not None:
            for q, out in ((in_qv, out_qv), (in_qc, out_qc), (in_qr, out_qr)):
                out[box] = _thomas(*_setup_tridiagonal(gamma, w, in_s[box] * q[box])) / out_s[box]


def implicit_vertical_advection_tendency(staggered, in_w, in_s, in_su, in_sv, tnd_s, tnd_su, tnd_sv, *,
                                         dt, gamma, origin, domain, in_qv=None, in_qc=None,
                                         in_qr=None, tnd_qv=None, tnd_qc=None, tnd_qr=None):
    """implicit_vertical_advection.py:L793-L919: the same solves, as tendencies (x_new - x) / dt."""
    box = tuple(slice(o, o + d) for o, d in zip(origin, domain))
    i, j, _ = box
    kb, ke = origin[2], origin[2] + domain[2]
    w = 0.5 * (in_w[i, j, kb:ke] + in_w[i, j, kb + 1:ke + 1]) if staggered else in_w[box]
    with np.errstate(all="ignore"):
        new_s = _thomas(*_setup_tridiagonal(gamma, w, in_s[box]))
        tnd_s[box] = (new_s - in_s[box]) / dt
        for src, out in ((in_su, tnd_su), (in_sv, tnd_sv)):
            out[box] = (_thomas(*_setup_tridiagonal(gamma, w, src[box])) - src[box]) / dt
        if in_qv is not None:
            for q, out in ((in_qv, tnd_qv), (in_qc, tnd_qc), (in_qr, tnd_qr)):
                out[box] = (_thomas(*_setup_tridiagonal(gamma, w, in_s[box] * q[box])) / new_s - q[box]) / dt
