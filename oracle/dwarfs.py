# -*- coding: utf-8 -*-
"""Oracle (test infrastructure): reusable numerics ("dwarfs"), rows K4, K6..K9, K12.

Follows
  src/tasmania/dwarfs/vertical_damping.py:L100-L111, subclasses/vertical_dampers/rayleigh.py:L90-L109
  src/tasmania/dwarfs/diagnostics.py:L175-L272 (momenta, velocity_x/y), L400-L450 (density, mass_fraction)
  src/tasmania/dwarfs/horizontal_diffusion.py:L89-L109, subclasses/horizontal_diffusers/{second,fourth}_order.py
  src/tasmania/dwarfs/horizontal_smoothing.py:L83-L94, subclasses/horizontal_smoothers/{first,second,third}_order.py
  (the one-dimensional ..._1dx / ..._1dy variants of both included)
  src/tasmania/framework/subclasses/stencil_definitions/{copy,math,algorithms}.py
"""
import math

import numpy as np


def _box(origin, domain):
    return tuple(slice(o, o + d) for o, d in zip(origin, domain))


def _sh(origin, domain, di=0, dj=0):
    """The (i, j, k) box shifted by (di, dj) points."""
    return (
        slice(origin[0] + di, origin[0] + domain[0] + di),
        slice(origin[1] + dj, origin[1] + domain[1] + dj),
        slice(origin[2], origin[2] + domain[2]),
    )


# ------------------------------------------------------------------ K6 Rayleigh damping
def rayleigh_coefficient(z_main, z_top, damp_depth, damp_max, nk):
    """vertical_damping.py:L100-L111.  ``z_main``: the nz main levels (decreasing with k),
    ``z_top``: z_on_interface_levels[0]; returns the rank-1 profile of length ``nk``."""
    nz = len(z_main)
    r = np.zeros(nk)
    if damp_depth > 0:
        z = np.concatenate((z_main, np.array([0]))) if nk == nz + 1 else np.asarray(z_main)
        za = z[damp_depth - 1]
        r = (z >= za) * damp_max * (1 - np.cos(math.pi * (z - za) / (z_top - za)))
    return r


def damping(phi_now, phi_new, phi_ref, rmat, out, dt, origin, domain):
    """rayleigh.py:L90-L109 -- out = new - (dt * R) * (now - ref), no mask."""
    b = _box(origin, domain)
    out[b] = phi_new[b] - dt * rmat[b] * (phi_now[b] - phi_ref[b])


# ------------------------------------------------------------------ K4 velocity / momenta
def velocity_x(d, du, u, origin, domain, staggering=True):
    """dwarfs/diagnostics.py:L219-L237."""
    c, m = _sh(origin, domain), _sh(origin, domain, di=-1)
    if staggering:
        u[c] = (du[m] + du[c]) / (d[m] + d[c])
    else:
        u[c] = du[c] / d[c]


def velocity_y(d, dv, v, origin, domain, staggering=True):
    """dwarfs/diagnostics.py:L254-L272."""
    c, m = _sh(origin, domain), _sh(origin, domain, dj=-1)
    if staggering:
        v[c] = (dv[m] + dv[c]) / (d[m] + d[c])
    else:
        v[c] = dv[c] / d[c]


def momenta(d, u, v, du, dv, origin, domain, staggering=True):
    """dwarfs/diagnostics.py:L175-L198."""
    c = _sh(origin, domain)
    if staggering:
        du[c] = 0.5 * d[c] * (u[c] + u[_sh(origin, domain, di=1)])
        dv[c] = 0.5 * d[c] * (v[c] + v[_sh(origin, domain, dj=1)])
    else:
        du[c] = d[c] * u[c]
        dv[c] = d[c] * v[c]


def get_velocity_components(nx, ny, nz, s, su, sv, u, v):
    """HorizontalVelocity.get_velocity_components, dwarfs/diagnostics.py:L125-L173 (staggered)."""
    velocity_x(s, su, u, (1, 0, 0), (nx - 1, ny, nz))
    velocity_y(s, sv, v, (0, 1, 0), (nx, ny - 1, nz))


# ------------------------------------------------------------------ K7 water constituents
def density(d, q, dq, origin, domain, clipping=True):
    """dwarfs/diagnostics.py:L400-L416."""
    b = _box(origin, domain)
    dq[b] = d[b] * q[b]
    if clipping:
        dq[b] = np.where(dq[b] > 0.0, dq[b], 0.0)


def mass_fraction(d, dq, q, origin, domain, clipping=True):
    """dwarfs/diagnostics.py:L434-L450."""
    b = _box(origin, domain)
    q[b] = dq[b] / d[b]
    if clipping:
        q[b] = np.where(q[b] > 0.0, q[b], 0.0)


# ------------------------------------------------------------------ K8 diffusion
def vertical_profile(coeff, coeff_max, damp_depth, nk):
    """gamma(k) of horizontal_diffusion.py:L91-L97 and horizontal_smoothing.py:L83-L89."""
    gamma = coeff * np.ones(nk)
    n = damp_depth
    if n > 0:
        pert = np.sin(0.5 * math.pi * (n - np.arange(0, n, dtype=float)) / n) ** 2
        gamma[:n] += (coeff_max - coeff) * pert
    return gamma


def diffusion(order, phi, gamma, out, dx, dy, overwrite, origin, domain):
    """second_order.py:L92-L106 / fourth_order.py:L92-L124 + set_output (generics.py:L38-L40).

    Like the reference's numpy definition, every k present is processed (origin[2] and
    domain[2] are ignored)."""
    i0, j0 = origin[0], origin[1]
    i1, j1 = i0 + domain[0], j0 + domain[1]

    def at(di, dj):
        return phi[i0 + di : i1 + di, j0 + dj : j1 + dj]

    g = gamma[i0:i1, j0:j1]
    if order == 2:
        tmp = g * (
            (at(-1, 0) - 2.0 * at(0, 0) + at(1, 0)) / (dx * dx)
            + (at(0, -1) - 2.0 * at(0, 0) + at(0, 1)) / (dy * dy)
        )
    elif order == 4:
        tmp = g * (
            (-at(-2, 0) + 16.0 * at(-1, 0) - 30.0 * at(0, 0) + 16.0 * at(1, 0) - at(2, 0))
            / (12.0 * dx * dx)
            + (-at(0, -2) + 16.0 * at(0, -1) - 30.0 * at(0, 0) + 16.0 * at(0, 1) - at(0, 2))
            / (12.0 * dy * dy)
        )
    else:
        raise ValueError(order)
    out[i0:i1, j0:j1] = tmp if overwrite else out[i0:i1, j0:j1] + tmp


# ------------------------------------------------------------------ K9 smoothing
def smoothing(order, phi, gamma, out, origin, domain):
    """first_order.py:L113-L126, second_order.py:L113-L139, third_order.py:L113-L150."""
    c = _sh(origin, domain)

    def at(di, dj):
        return phi[_sh(origin, domain, di, dj)]

    g = gamma[c]
    if order == 1:
        out[c] = (1.0 - g) * phi[c] + 0.25 * g * (at(-1, 0) + at(1, 0) + at(0, -1) + at(0, 1))
    elif order == 2:
        out[c] = (1.0 - 0.75 * g) * phi[c] + 0.0625 * g * (
            -at(-2, 0) + 4.0 * at(-1, 0) - at(2, 0) + 4.0 * at(1, 0)
            - at(0, -2) + 4.0 * at(0, -1) - at(0, 2) + 4.0 * at(0, 1)
        )
    elif order == 3:
        out[c] = (1.0 - 0.625 * g) * phi[c] + 0.015625 * g * (
            at(-3, 0) - 6.0 * at(-2, 0) + 15.0 * at(-1, 0)
            + at(3, 0) - 6.0 * at(2, 0) + 15.0 * at(1, 0)
            + at(0, -3) - 6.0 * at(0, -2) + 15.0 * at(0, -1)
            + at(0, 3) - 6.0 * at(0, 2) + 15.0 * at(0, 1)
        )
    else:
        raise ValueError(order)


def horizontal_smoothing(order, phi, gamma, out, shape=None):
    """HorizontalSmoothing.__call__ (e.g. first_order.py:L60-L110): smoothing on the
    interior + four rim copies."""
    nx, ny, nz = shape or phi.shape
    nb = order
    smoothing(order, phi, gamma, out, (nb, nb, 0), (nx - 2 * nb, ny - 2 * nb, nz))
    for o, d in (
        ((0, 0, 0), (nb, ny, nz)),
        ((nx - nb, 0, 0), (nb, ny, nz)),
        ((nb, 0, 0), (nx - 2 * nb, nb, nz)),
        ((nb, ny - nb, 0), (nx - 2 * nb, nb, nz)),
    ):
        copy(phi, out, o, d)


def hyperdiffusion(phi, out, alpha, origin, domain):
    """The class-less ``diffusion`` stencil, framework/subclasses/stencil_definitions/
    diffusion.py:L31-L55: phi + alpha * (differences of the fluxes of the bilaplacian)."""
    ib, jb, kb = origin
    ie, je, ke = (o + d for o, d in zip(origin, domain))
    k = slice(kb, ke)

    def sh(lo_i, hi_i, lo_j, hi_j):
        return phi[ib + lo_i : ie + hi_i, jb + lo_j : je + hi_j, k]

    lap = -4 * sh(-2, 2, -2, 2) + sh(-3, 1, -2, 2) + sh(-1, 3, -2, 2) + sh(-2, 2, -3, 1) + sh(-2, 2, -1, 3)
    bilap = -4 * lap[1:-1, 1:-1] + lap[:-2, 1:-1] + lap[2:, 1:-1] + lap[1:-1, :-2] + lap[1:-1, 2:]
    flux_x = bilap[1:, 1:-1] - bilap[:-1, 1:-1]
    flux_y = bilap[1:-1, 1:] - bilap[1:-1, :-1]
    out[ib:ie, jb:je, k] = phi[ib:ie, jb:je, k] + alpha * (
        flux_x[1:, :] - flux_x[:-1, :] + flux_y[:, 1:] - flux_y[:, :-1]
    )


# ------------------------------------------------------------------ K8 / K9, 1-D variants
def diffusion_1d(order, axis, phi, gamma, out, h, overwrite, origin, domain):
    """SecondOrder1DX / 1DY (horizontal_diffusers/second_order.py:L210-L219, L321-L330) and
    FourthOrder1DX / 1DY (fourth_order.py:L256-L277, L393-L414): the stencil along one axis,
    ``gamma * (...) / (h * h)`` with the product first; every k present is processed."""
    i0, j0 = origin[0], origin[1]
    i1, j1 = i0 + domain[0], j0 + domain[1]

    def at(m):
        di, dj = (m, 0) if axis == 0 else (0, m)
        return phi[i0 + di : i1 + di, j0 + dj : j1 + dj]

    g = gamma[i0:i1, j0:j1]
    if order == 2:
        tmp = g * (at(-1) - 2.0 * at(0) + at(1)) / (h * h)
    elif order == 4:
        tmp = g * (-at(-2) + 16.0 * at(-1) - 30.0 * at(0) + 16.0 * at(1) - at(2)) / (12.0 * h * h)
    else:
        raise ValueError(order)
    out[i0:i1, j0:j1] = tmp if overwrite else out[i0:i1, j0:j1] + tmp


def smoothing_1d(order, axis, phi, gamma, out, origin, domain):
    """FirstOrder1DX / 1DY (horizontal_smoothers/first_order.py:L209-L218, L301-L310),
    SecondOrder1DX / 1DY (second_order.py:L231-L247, L335-L351), ThirdOrder1DX / 1DY
    (third_order.py:L243-L263, L353-L373)."""
    c = _sh(origin, domain)

    def at(m):
        return phi[_sh(origin, domain, m, 0) if axis == 0 else _sh(origin, domain, 0, m)]

    g = gamma[c]
    if order == 1:
        out[c] = (1.0 - 0.5 * g) * phi[c] + 0.25 * g * (at(-1) + at(1))
    elif order == 2:
        out[c] = (1.0 - 0.375 * g) * phi[c] + 0.0625 * g * (
            -at(-2) + 4.0 * at(-1) - at(2) + 4.0 * at(1)
        )
    elif order == 3:
        out[c] = (1.0 - 0.3125 * g) * phi[c] + 0.015625 * g * (
            at(-3) - 6.0 * at(-2) + 15.0 * at(-1) + at(3) - 6.0 * at(2) + 15.0 * at(1)
        )
    else:
        raise ValueError(order)


def horizontal_smoothing_1d(order, axis, phi, gamma, out, shape=None):
    """The 1-D smoothers' ``__call__`` (e.g. first_order.py:L170-L205, L262-L297): smoothing
    on the interior along the axis + the two rim copies across it."""
    nx, ny, nz = shape or phi.shape
    nb = order
    if axis == 0:
        smoothing_1d(order, 0, phi, gamma, out, (nb, 0, 0), (nx - 2 * nb, ny, nz))
        copy(phi, out, (0, 0, 0), (nb, ny, nz))
        copy(phi, out, (nx - nb, 0, 0), (nb, ny, nz))
    else:
        smoothing_1d(order, 1, phi, gamma, out, (0, nb, 0), (nx, ny - 2 * nb, nz))
        copy(phi, out, (0, 0, 0), (nx, nb, nz))
        copy(phi, out, (0, ny - nb, 0), (nx, nb, nz))


# ------------------------------------------------------------------ K12 elementwise
def copy(src, dst, origin, domain):
    """stencil_definitions/copy.py:L30-L34."""
    b = _box(origin, domain)
    dst[b] = src[b]


def copychange(src, dst, origin, domain):
    """stencil_definitions/copy.py:L37-L41."""
    b = _box(origin, domain)
    dst[b] = -src[b]


ELEMENTWISE = {
    # name: (n_inputs, has_f, fn(inputs..., f) -> out)   math.py:L32-L124
    "abs": lambda a: np.abs(a),
    "add": lambda a, b: a + b,
    "addsub": lambda a, b, c: a + b - c,
    "clip": lambda a: np.where(a > 0.0, a, 0.0),
    "fma": lambda a, b, f: a + f * b,
    "mul": lambda a, b: a * b,
    "scale": lambda a, f: f * a,
    "sub": lambda a, b: a - b,
    "sts_rk2_0": lambda a, prv, tnd, dt: 0.5 * (a + prv + dt * tnd),
    "sts_rk3ws_0": lambda a, prv, tnd, dt: (2.0 * a + prv + dt * tnd) / 3.0,
}
