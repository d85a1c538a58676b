# -*- coding: utf-8 -*-
"""Oracle (test infrastructure): the isentropic dynamical core, rows K1..K3 of SURVEY.md
section 8a, and the stage orchestration around them.

Follows
  src/tasmania/isentropic/dynamics/subclasses/prognostics/utils.py:L43-L204       (K1, K2)
  src/tasmania/isentropic/dynamics/diagnostics.py:L319-L360, L408-L438, L472-L503, L540-L570 (K3)
  src/tasmania/isentropic/dynamics/subclasses/prognostics/rk3ws_si.py:L105-L234   (RK3WS-SI stage)
  src/tasmania/isentropic/dynamics/subclasses/prognostics/forward_euler_si.py:L93-L196
  src/tasmania/isentropic/dynamics/dycore.py:L641-L843                            (stage_array_call)
  src/tasmania/framework/dycore.py:L455-L462                                      (stage chaining)
"""
from copy import deepcopy

import numpy as np

from oracle import dwarfs
from oracle.fluxes import EXTENT, face_flux

# default physical constants, isentropic/dynamics/diagnostics.py:L56-L63
CONSTANTS = {"pref": 1.0e5, "rd": 287.05, "g": 9.80665, "cp": 1004.0}

MFWV = "mass_fraction_of_water_vapor_in_air"
MFCW = "mass_fraction_of_cloud_liquid_water_in_air"
MFPW = "mass_fraction_of_precipitation_water_in_air"
S, SU, SV = "air_isentropic_density", "x_momentum_isentropic", "y_momentum_isentropic"
U, V = "x_velocity_at_u_locations", "y_velocity_at_v_locations"
MTG = "montgomery_potential"
SQV = "isentropic_density_of_water_vapor"
SQC = "isentropic_density_of_cloud_liquid_water"
SQR = "isentropic_density_of_precipitation_water"


# ------------------------------------------------------------------ K1 / K2
def _divergence(scheme, u, v, phi, origin, domain, dx, dy):
    """(Fx[i+1/2]-Fx[i-1/2])/dx + (Fy[j+1/2]-Fy[j-1/2])/dy over the box, utils.py:L77-L103."""
    e = EXTENT[scheme]
    i0, j0, k0 = origin
    ni, nj, nk = domain
    i, j, k = slice(i0, i0 + ni), slice(j0, j0 + nj), slice(k0, k0 + nk)
    ip, im = slice(i0 - e + 1, i0 - e + ni + 1), slice(i0 - e, i0 - e + ni)
    jp, jm = slice(j0 - e + 1, j0 - e + nj + 1), slice(j0 - e, j0 - e + nj)
    fx = face_flux(scheme, u, phi, 0)
    fy = face_flux(scheme, v, phi, 1)
    return (fx[ip, j, k] - fx[im, j, k]) / dx + (fy[i, jp, k] - fy[i, jm, k]) / dy


def step_forward_euler(
    scheme, s_now, s_int, s_new, u_int, v_int, *, dt, dx, dy, origin, domain,
    s_tnd=None, moist=False, sq_now=(), sq_int=(), sq_new=(), q_tnd=(None, None, None),
):
    """K1, utils.py:L43-L134."""
    i0, j0, k0 = origin
    b = (slice(i0, i0 + domain[0]), slice(j0, j0 + domain[1]), slice(k0, k0 + domain[2]))
    s_new[b] = s_now[b] - dt * (
        _divergence(scheme, u_int, v_int, s_int, origin, domain, dx, dy)
        - (s_tnd[b] if s_tnd is not None else 0.0)
    )
    if moist:
        for now, int_, new, tnd in zip(sq_now, sq_int, sq_new, q_tnd):
            new[b] = now[b] - dt * (
                _divergence(scheme, u_int, v_int, int_, origin, domain, dx, dy)
                - (s_int[b] * tnd[b] if tnd is not None else 0.0)
            )


def step_forward_euler_momentum(
    scheme, s_now, s_new, u_int, v_int, su_now, su_int, su_new, sv_now, sv_int, sv_new,
    mtg_now, mtg_new, *, dt, dx, dy, eps, origin, domain, su_tnd=None, sv_tnd=None,
):
    """K2, utils.py:L137-L204."""
    i0, j0, k0 = origin
    ni, nj, nk = domain
    i, j, k = slice(i0, i0 + ni), slice(j0, j0 + nj), slice(k0, k0 + nk)
    im1, ip1 = slice(i0 - 1, i0 + ni - 1), slice(i0 + 1, i0 + ni + 1)
    jm1, jp1 = slice(j0 - 1, j0 + nj - 1), slice(j0 + 1, j0 + nj + 1)
    b = (i, j, k)
    su_new[b] = su_now[b] - dt * (
        _divergence(scheme, u_int, v_int, su_int, origin, domain, dx, dy)
        + (1.0 - eps) * s_now[b] * (mtg_now[ip1, j, k] - mtg_now[im1, j, k]) / (2.0 * dx)
        + eps * s_new[b] * (mtg_new[ip1, j, k] - mtg_new[im1, j, k]) / (2.0 * dx)
        - (su_tnd[b] if su_tnd is not None else 0.0)
    )
    sv_new[b] = sv_now[b] - dt * (
        _divergence(scheme, u_int, v_int, sv_int, origin, domain, dx, dy)
        + (1.0 - eps) * s_now[b] * (mtg_now[i, jp1, k] - mtg_now[i, jm1, k]) / (2.0 * dy)
        + eps * s_new[b] * (mtg_new[i, jp1, k] - mtg_new[i, jm1, k]) / (2.0 * dy)
        - (sv_tnd[b] if sv_tnd is not None else 0.0)
    )


# ------------------------------------------------------------------ K3 column scans
def montgomery(hs, s, mtg, *, dz, pt, theta_s, origin, domain, constants=CONSTANTS):
    """diagnostics.py:L408-L438.  ``hs`` is a 3-D storage holding the topography at level
    kstop-1; sequential top-down pressure scan, bottom-up Montgomery scan."""
    g, cp, pref, rd = (constants[n] for n in ("g", "cp", "pref", "rd"))
    i = slice(origin[0], origin[0] + domain[0])
    j = slice(origin[1], origin[1] + domain[1])
    k0, k1 = origin[2], origin[2] + domain[2]
    p = deepcopy(s)
    p[i, j, k0] = pt
    for k in range(k0 + 1, k1):
        p[i, j, k] = p[i, j, k - 1] + g * dz * s[i, j, k - 1]
    exn = cp * (p / pref) ** (rd / cp)
    mtg_s = theta_s * exn[i, j, k1 - 1] + g * hs[i, j, k1 - 1]
    mtg[i, j, k1 - 2] = mtg_s + 0.5 * dz * exn[i, j, k1 - 1]
    for k in range(k1 - 3, k0 - 1, -1):
        mtg[i, j, k] = mtg[i, j, k + 1] + dz * exn[i, j, k + 1]


def diagnostic_variables(theta, hs, s, p, exn, mtg, h, *, dz, pt, origin, domain,
                         constants=CONSTANTS):
    """diagnostics.py:L319-L360 -- p, exn, mtg, h from s."""
    g, cp, pref, rd = (constants[n] for n in ("g", "cp", "pref", "rd"))
    i = slice(origin[0], origin[0] + domain[0])
    j = slice(origin[1], origin[1] + domain[1])
    k0, k1 = origin[2], origin[2] + domain[2]
    p[i, j, k0] = pt
    for k in range(k0 + 1, k1):
        p[i, j, k] = p[i, j, k - 1] + g * dz * s[i, j, k - 1]
    exn[i, j, k0:k1] = cp * (p[i, j, k0:k1] / pref) ** (rd / cp)
    mtg_s = theta[i, j, k1 - 1] * exn[i, j, k1 - 1] + g * hs[i, j, k1 - 1]
    mtg[i, j, k1 - 2] = mtg_s + 0.5 * dz * exn[i, j, k1 - 1]
    for k in range(k1 - 3, k0 - 1, -1):
        mtg[i, j, k] = mtg[i, j, k + 1] + dz * exn[i, j, k + 1]
    h[i, j, k1 - 1] = hs[i, j, k1 - 1]
    for k in range(k1 - 2, k0 - 1, -1):
        h[i, j, k] = h[i, j, k + 1] - rd * (
            theta[i, j, k] * exn[i, j, k] + theta[i, j, k + 1] * exn[i, j, k + 1]
        ) * (p[i, j, k] - p[i, j, k + 1]) / (cp * g * (p[i, j, k] + p[i, j, k + 1]))


def height(theta, hs, s, h, *, dz, pt, origin, domain, constants=CONSTANTS):
    """diagnostics.py:L472-L503."""
    g, cp, pref, rd = (constants[n] for n in ("g", "cp", "pref", "rd"))
    i = slice(origin[0], origin[0] + domain[0])
    j = slice(origin[1], origin[1] + domain[1])
    k0, k1 = origin[2], origin[2] + domain[2]
    p = deepcopy(s)
    p[i, j, k0] = pt
    for k in range(k0 + 1, k1):
        p[i, j, k] = p[i, j, k - 1] + g * dz * s[i, j, k - 1]
    exn = cp * (p / pref) ** (rd / cp)
    h[i, j, k1 - 1] = hs[i, j, k1 - 1]
    for k in range(k1 - 2, k0 - 1, -1):
        h[i, j, k] = h[i, j, k + 1] - rd * (
            theta[i, j, k] * exn[i, j, k] + theta[i, j, k + 1] * exn[i, j, k + 1]
        ) * (p[i, j, k] - p[i, j, k + 1]) / (cp * g * (p[i, j, k] + p[i, j, k + 1]))


def density_and_temperature(theta, s, exn, h, rho, t, *, origin, domain, constants=CONSTANTS):
    """diagnostics.py:L540-L570."""
    cp = constants["cp"]
    i = slice(origin[0], origin[0] + domain[0])
    j = slice(origin[1], origin[1] + domain[1])
    k = slice(origin[2], origin[2] + domain[2])
    kp1 = slice(origin[2] + 1, origin[2] + domain[2] + 1)
    rho[i, j, k] = s[i, j, k] * (theta[i, j, k] - theta[i, j, kp1]) / (h[i, j, k] - h[i, j, kp1])
    t[i, j, k] = 0.5 / cp * (theta[i, j, k] * exn[i, j, k] + theta[i, j, kp1] * exn[i, j, kp1])


# ------------------------------------------------------------------ orchestration
class Grid:
    """The few grid numbers the numerics need (numerical grid)."""

    def __init__(self, nx, ny, nz, dx, dy, dz, z_on_interface_levels, z_main=None):
        self.nx, self.ny, self.nz = nx, ny, nz
        self.dx, self.dy, self.dz = dx, dy, dz
        self.z_hl = np.asarray(z_on_interface_levels, dtype=float)
        self.z = (
            np.asarray(z_main, dtype=float) if z_main is not None
            else 0.5 * (self.z_hl[:-1] + self.z_hl[1:])
        )


class IsentropicDycore:
    """Raw-array restatement of ``IsentropicDynamicalCore`` (dry and moist), no smoothing.

    ``hb``: oracle.boundary.Relaxed / Periodic with ``reference_state`` set (raw arrays);
    ``topo``: callable returning the current 2-D topography [m] of shape (nx, ny).
    """

    def __init__(self, grid, hb, topo, *, moist=False, scheme="rk3ws_si",
                 flux="fifth_order_upwind", pt=0.0, eps=0.5, damp=True, damp_at_every_stage=True,
                 damp_depth=15, damp_max=0.0002, constants=CONSTANTS, shape=None):
        self.g, self.hb, self.topo = grid, hb, topo
        self.moist, self.scheme, self.flux = moist, scheme, flux
        self.pt, self.eps = pt, eps
        self.damp, self.damp_every = damp, damp_at_every_stage
        self.constants = constants
        nx, ny, nz = grid.nx, grid.ny, grid.nz
        self.shape = shape or (nx + 1, ny + 1, nz + 1)
        assert hb.nb >= EXTENT[flux] and nx >= 2 * hb.nb + 1 and ny >= 2 * hb.nb + 1
        self.stages = {"forward_euler_si": 1, "rk3ws_si": 3}[scheme]
        if damp:
            r = dwarfs.rayleigh_coefficient(grid.z, grid.z_hl[0], damp_depth, damp_max, self.shape[2])
            self.rmat = np.zeros(self.shape)
            self.rmat[...] = r[None, None, :]
        self.mtg_new = np.zeros(self.shape)
        self._topo3d = np.zeros(self.shape)
        self._stage_states = None

    # ---- prognostic stage: rk3ws_si.py:L105-L234 / forward_euler_si.py:L93-L196
    def _dts(self, stage, timestep):
        if self.scheme == "forward_euler_si":
            return timestep, timestep
        if stage == 0:
            return timestep / 3.0, timestep / 3.0  # timedelta arithmetic (microsecond rounding)
        if stage == 1:
            return timestep / 6.0, 0.5 * timestep
        return 0.5 * timestep, timestep

    def _prognostic(self, stage, timestep, state, tendencies, out):
        g, nb = self.g, self.hb.nb
        nx, ny, nz = g.nx, g.ny, g.nz
        dtr, dt = self._dts(stage, timestep)
        dt = dt.total_seconds()
        if stage == 0:
            self._now = {n: state[n] for n in (S, MTG, SU, SV)}
            if self.moist:
                self._now.update({n: state[n] for n in (SQV, SQC, SQR)})
        origin, domain = (nb, nb, 0), (nx - 2 * nb, ny - 2 * nb, nz)
        kw = {}
        if self.moist:
            kw = dict(
                moist=True,
                sq_now=[self._now[n] for n in (SQV, SQC, SQR)],
                sq_int=[state[n] for n in (SQV, SQC, SQR)],
                sq_new=[out[n] for n in (SQV, SQC, SQR)],
                q_tnd=[tendencies.get(n) for n in (MFWV, MFCW, MFPW)],
            )
        step_forward_euler(
            self.flux, self._now[S], state[S], out[S], state[U], state[V],
            dt=dt, dx=g.dx, dy=g.dy, origin=origin, domain=domain, s_tnd=tendencies.get(S), **kw,
        )
        self.hb.enforce_field(out[S], S)
        self._topo3d[:nx, :ny, nz] = self.topo()
        montgomery(self._topo3d, out[S], self.mtg_new, dz=g.dz, pt=self.pt, theta_s=g.z_hl[-1],
                   origin=(0, 0, 0), domain=(nx, ny, nz + 1), constants=self.constants)
        step_forward_euler_momentum(
            self.flux, self._now[S], out[S], state[U], state[V],
            self._now[SU], state[SU], out[SU], self._now[SV], state[SV], out[SV],
            self._now[MTG], self.mtg_new, dt=dt, dx=g.dx, dy=g.dy, eps=self.eps,
            origin=origin, domain=domain, su_tnd=tendencies.get(SU), sv_tnd=tendencies.get(SV),
        )
        out["time"] = state["time"] + dtr

    # ---- one stage: dycore.py:L641-L843
    def stage_array_call(self, stage, state, tendencies, timestep, out):
        g, hb = self.g, self.hb
        nx, ny, nz = g.nx, g.ny, g.nz
        if self.damp and stage == 0:
            self._ref = {n: hb.reference_state[n] for n in (S, SU, SV)}
            self._dnow = {n: state[n] for n in (S, SU, SV)}
        if self.moist:
            box = ((0, 0, 0), (nx, ny, nz))
            tag = "now" if stage == 0 else "int"
            for qn, sqn in ((MFWV, SQV), (MFCW, SQC), (MFPW, SQR)):
                buf = self._sq.setdefault((tag, sqn), np.zeros(self.shape))
                dwarfs.density(state[S], state[qn], buf, *box)
                state[sqn] = buf
                out[sqn] = self._sq.setdefault(("new", sqn), np.zeros(self.shape))
        self._prognostic(stage, timestep, state, tendencies, out)
        if self.moist:
            for qn, sqn in ((MFWV, SQV), (MFCW, SQC), (MFPW, SQR)):
                dwarfs.mass_fraction(out[S], out.pop(sqn), out[qn], (0, 0, 0), (nx, ny, nz))
        names = (S, SU, U, SV, V) + ((MFWV, MFCW, MFPW) if self.moist else ())
        hb.enforce_raw(out, names)
        if self.damp and (self.damp_every or stage == self.stages - 1):
            full = ((0, 0, 0), self.shape)
            dtf = timestep.total_seconds()
            for n in (S, SU, SV):
                dwarfs.damping(self._dnow[n], out[n], self._ref[n], self.rmat, out[n], dtf, *full)
        dwarfs.get_velocity_components(nx, ny, nz, out[S], out[SU], out[SV], out[U], out[V])
        hb.set_outermost_layers_x(out[U], U)
        hb.set_outermost_layers_y(out[V], V)

    _sq = None

    def allocate_outputs(self):
        names = (S, SU, U, SV, V) + ((MFWV, MFCW, MFPW) if self.moist else ())
        return {n: np.zeros(self.shape) for n in names}

    # ---- one time step: framework/dycore.py:L383-L462
    def __call__(self, state, tendencies, timestep, out_state=None):
        if self._sq is None:
            self._sq = {}
        if self._stage_states is None:
            self._stage_states = [self.allocate_outputs() for _ in range(self.stages - 1)]
        out_state = out_state if out_state is not None else self.allocate_outputs()
        outs = self._stage_states + [out_state]
        cur = state
        for stage in range(self.stages):
            # the stage sees a shallow copy: the moist path adds sq* entries to it
            self.stage_array_call(stage, dict(cur), tendencies or {}, timestep, outs[stage])
            cur = outs[stage]
        out_state["time"] = state["time"] + timestep
        return out_state


def refresh_diagnostics(grid, topo2d, s, pt, p, exn, mtg, h, constants=CONSTANTS):
    """IsentropicDiagnostics.get_diagnostic_variables, diagnostics.py:L140-L193: the role
    the ``dv`` component plays after each dycore step (driver_namelist_sus.py:L188-L199)."""
    nx, ny, nz = grid.nx, grid.ny, grid.nz
    shape = s.shape
    theta = np.zeros(shape)
    theta[:nx, :ny, : nz + 1] = grid.z_hl[None, None, :]
    topo = np.zeros(shape)
    topo[:nx, :ny, nz] = topo2d
    diagnostic_variables(theta, topo, s, p, exn, mtg, h, dz=grid.dz, pt=pt,
                         origin=(0, 0, 0), domain=(nx, ny, nz + 1), constants=constants)
