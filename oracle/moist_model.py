# -*- coding: utf-8 -*-
"""ORACLE (test infrastructure only, never imported by tasmania_b200): the moist isentropic model of
BASELINE.json configs[2] -- dynamical core + sequential-update-splitting physics -- as one
straight-line numpy function per time step.

It restates, for the namelist of the reference's own benchmark driver
(drivers/benchmarking/isentropic_moist/namelist_sus.py, driver_namelist_sus.py:L184-L512), what
the reference's coupling classes do at the raw-array level:

  SequentialUpdateSplitting.__call__   src/tasmania/framework/sequential_update_splitting.py:L161-L194
  ConcurrentCoupling._call_serial      src/tasmania/framework/concurrent_coupling.py:L314-L374
  overwrite / accumulate of tendencies src/tasmania/framework/concurrent_coupling_utils.py:L72-L83
  TendencyStepper forward_euler / rk2 / rk3ws
                                       src/tasmania/framework/subclasses/tendency_steppers/*.py
  DataArrayDictOperator.fma            src/tasmania/utils/xarrayx.py:L688-L740 (stencil math.py:L59-L63,
                                       on the whole storage)
  promoters t2d / d2t                  src/tasmania/framework/promoter.py:L161-L176, L290-L305,
                                       src/tasmania/isentropic/utils.py:L27-L62
  component array_calls                src/tasmania/isentropic/physics/{diagnostics,coriolis,
                                       horizontal_smoothing,turbulence,vertical_advection}.py,
                                       src/tasmania/physics/microphysics/{kessler,utils}.py

Deliberately written without coupler classes (the product mirrors the reference's class structure
in tasmania_b200/coupling.py; this file spells the resulting sequence out), so that the two are
independent statements of the same thing.  The stencils it calls are the oracle functions pinned
bit for bit on the reference's numpy definitions (tests/test_oracle_golden.py), and
``tendency_step`` is pinned bit for bit on the reference's own ForwardEuler / RK2 / RK3WS ``_call``
methods executed in place (tests/test_coupling_reference.py, which also runs the reference's
ConcurrentCoupling._call_serial, promoters and SequentialUpdateSplitting.__call__ against the b200
mirrors).  ``physics`` as a whole is pinned bit for bit on the reference's own component objects
chained by the reference's own couplers and steppers (tests/test_moist_physics_reference.py).
UNPINNED BY EXECUTION: only the component list and order, taken by reading from the benchmark
driver (a script over sympl objects that cannot run here).
"""
from __future__ import annotations

import numpy as np

from oracle import dwarfs, isentropic as oi, isentropic_physics as op, microphysics as om

S, SU, SV, U, V, MTG = oi.S, oi.SU, oi.SV, oi.U, oi.V, oi.MTG
QV, QC, QR = oi.MFWV, oi.MFCW, oi.MFPW
P = "air_pressure_on_interface_levels"
EXN = "exner_function_on_interface_levels"
H = "height_on_interface_levels"
RHO, T = "air_density", "air_temperature"
W = "tendency_of_air_potential_temperature"
THETA = "air_potential_temperature"
VT = "raindrop_fall_velocity"
PREC, ACCPREC = "precipitation", "accumulated_precipitation"

STAGE_FACTORS = {  # forward_euler.py:L57-L72, rk2.py:L60-L118, rk3ws.py:L60-L157
    "forward_euler": (1.0,),
    "rk2": (0.5, 1.0),
    "rk3ws": (1.0 / 3.0, 0.5, 1.0),
}


def tendency_step(scheme, state, tendency_fn, dt_s):
    """One TendencyStepper call: returns (diagnostics of the FIRST stage, stepped fields).

    ``tendency_fn(state) -> (tendencies, diagnostics)`` plays ``get_increment``; every stage
    restarts from ``state`` (``fma(state, increment, c * dt)``), the intermediate state is ``state``
    with the stepped fields replaced."""
    first, cur, out = None, state, {}
    for n, c in enumerate(STAGE_FACTORS[scheme]):
        tnd, diag = tendency_fn(cur)
        if n == 0:
            first = diag
        f = c * dt_s
        out = {k: state[k] + f * tnd[k] for k in tnd if k in state}
        cur = dict(state)
        cur.update(out)
    return first, out


class MoistIsentropicModel:
    """``step(state, dt)`` = one pass of the benchmark loop body, driver_namelist_sus.py:L490-L512."""

    def __init__(self, grid, hb, topo, pt, *, flux="fifth_order_upwind", eps=0.5, damp_depth=15,
                 damp_max=5e-4, physics_scheme="rk2", coriolis_parameter=1e-4,
                 smooth_order=2, smooth_coeff=1.0, smooth_coeff_max=1.0, smooth_damp_depth=0,
                 smagorinsky_constant=0.18, vertical_flux="third_order_upwind",
                 sedimentation_order=2, autoconversion_threshold=1e-4, autoconversion_rate=1e-3,
                 collection_rate=2.2, saturation_rate=0.025, shape=None):
        self.g, self.hb, self.topo, self.pt = grid, hb, topo, pt
        nx, ny, nz = grid.nx, grid.ny, grid.nz
        self.shape = tuple(shape or (nx + 1, ny + 1, nz + 1))
        self.dycore = oi.IsentropicDycore(
            grid, hb, topo, moist=True, scheme="rk3ws_si", flux=flux, pt=pt, eps=eps, damp=True,
            damp_at_every_stage=False, damp_depth=damp_depth, damp_max=damp_max, shape=self.shape)
        self.ptis = physics_scheme
        self.f = coriolis_parameter
        self.smooth_order = smooth_order
        gamma = dwarfs.vertical_profile(smooth_coeff, smooth_coeff_max, smooth_damp_depth, self.shape[2])
        self.gamma = np.zeros(self.shape)
        self.gamma[...] = gamma[None, None, :]
        self.cs = smagorinsky_constant
        self.vflux, self.sed_order = vertical_flux, sedimentation_order
        self.a, self.k1, self.k2, self.sr = (autoconversion_threshold, autoconversion_rate,
                                             collection_rate, saturation_rate)
        self.theta = np.zeros(self.shape)
        self.theta[:nx, :ny, : nz + 1] = grid.z_hl[None, None, :]
        self.nstep = 0

    def z(self, shape=None):
        return np.zeros(shape or self.shape)

    # ---------------------------------------------------------------- tendency providers
    def _coriolis(self, st):  # coriolis.py:L139-L164
        g, nb = self.g, self.hb.nb
        t = {SU: self.z(), SV: self.z()}
        op.coriolis(st[SU], st[SV], t[SU], t[SV], f=self.f, ow_tnd_su=True, ow_tnd_sv=True,
                    origin=(nb, nb, 0), domain=(g.nx - 2 * nb, g.ny - 2 * nb, g.nz))
        return t, {}

    def _smagorinsky(self, st):  # isentropic/physics/turbulence.py:L99-L125
        g, nb = self.g, max(2, self.hb.nb)
        t = {SU: self.z(), SV: self.z()}
        op.smagorinsky(st[SU], st[SV], t[SU], t[SV], dx=g.dx, dy=g.dy, cs=self.cs,
                       ow_out_u_tnd=True, ow_out_v_tnd=True, origin=(nb, nb, 0),
                       domain=(g.nx - 2 * nb, g.ny - 2 * nb, g.nz), in_s=st[S])
        return t, {}

    def _box(self):
        g = self.g
        return dict(origin=(0, 0, 0), domain=(g.nx, g.ny, g.nz))

    def _promote(self, theta_tnd):  # promoter.py:L290-L305: copy on the grid box
        g = self.g
        w = self.z()
        w[: g.nx, : g.ny, : g.nz] = theta_tnd[: g.nx, : g.ny, : g.nz]
        return w

    def _kessler(self, st):  # ConcurrentCoupling(ke, t2d), driver L286-L318
        t = {QV: self.z(), QC: self.z(), QR: self.z(), THETA: self.z()}
        om.kessler(st[RHO], st[P], st[T], st[EXN], st[QC], st[QR], st[QV], t[QC], t[QR], t[QV],
                   t[THETA], a=self.a, k1=self.k1, k2=self.k2, ow_out_qc_tnd=True,
                   ow_out_qr_tnd=True, ow_out_qv_tnd=True, ow_out_theta_tnd=True, **self._box())
        return t, {W: self._promote(t[THETA])}

    def _saturation(self, st):  # ConcurrentCoupling(d2t, sa, t2d), driver L320-L366
        g = self.g
        t = {QV: self.z(), QC: self.z(), THETA: self.z()}
        t[THETA][: g.nx, : g.ny, : g.nz] = st[W][: g.nx, : g.ny, : g.nz]  # d2t, promoter.py:L161-L176
        om.saturation_prognostic(st[P], st[T], st[EXN], st[QV], st[QC], t[QV], t[QC], t[THETA],
                                 sr=self.sr, ow_tnd_qv=True, ow_tnd_qc=True, ow_tnd_theta=False,
                                 **self._box())
        return t, {W: self._promote(t[THETA])}

    def _vertical_advection(self, st):  # vertical_advection.py:L216-L269, w on main levels
        names = (S, SU, SV, QV, QC, QR)
        t = {n: self.z() for n in names}
        op.vertical_advection(
            self.vflux, False, st[W], st[S], st[SU], st[SV], t[S], t[SU], t[SV], dz=self.g.dz,
            ow_out_s=True, ow_out_su=True, ow_out_sv=True, in_qv=st[QV], in_qc=st[QC], in_qr=st[QR],
            out_qv=t[QV], out_qc=t[QC], out_qr=t[QR], **self._box())
        return t, {}

    def _fall_velocity(self, st):  # kessler.py:L1167-L1181
        g = self.g
        rho_s = self.z()
        rho_s[: g.nx, : g.ny, : g.nz] = st[RHO][: g.nx, : g.ny, g.nz - 1 : g.nz]
        vt = self.z()
        om.fall_velocity(st[RHO], rho_s, st[QR], vt, **self._box())
        return vt

    def _sedimentation(self, st):  # ConcurrentCoupling(rfv, sd), driver L418-L452
        vt = self._fall_velocity(st)
        t = {QR: self.z()}
        om.sedimentation(st[RHO], st[H], st[QR], vt, t[QR], ow_out_tnd_qr=True,
                         order=self.sed_order, **self._box())
        return t, {VT: vt}

    def _precipitation(self, st, dt_s):  # ConcurrentCoupling(rfv, ap), driver L454-L476
        g = self.g
        vt = self._fall_velocity(st)
        shape2d = (self.shape[0], self.shape[1], 1)
        prec, acc = self.z(shape2d), self.z(shape2d)
        k = slice(g.nz - 1, g.nz)
        om.accumulated_precipitation(st[RHO][:, :, k], st[QR][:, :, k], vt[:, :, k], st[ACCPREC],
                                     prec, acc, dt=dt_s, origin=(0, 0, 0), domain=(g.nx, g.ny, 1))
        return {}, {VT: vt, PREC: prec, ACCPREC: acc}

    # ---------------------------------------------------------------- physics
    def physics(self, state, dt):
        """SequentialUpdateSplitting over the ten entries of driver_namelist_sus.py:L184-L479;
        ``state`` is updated in place."""
        g, hb = self.g, self.hb
        nx, ny, nz = g.nx, g.ny, g.nz
        dt_s = dt.total_seconds()
        # 1. dv: IsentropicDiagnostics(moist=True), isentropic/physics/diagnostics.py:L175-L196
        p, exn, mtg, h = self.z(), self.z(), self.z(), self.z()
        oi.refresh_diagnostics(g, self.topo(), state[S], self.pt, p, exn, mtg, h)
        rho, t = self.z(), self.z()
        oi.density_and_temperature(self.theta, state[S], exn, h, rho, t, origin=(0, 0, 0),
                                   domain=(nx, ny, nz))
        state.update({P: p, EXN: exn, MTG: mtg, H: h, RHO: rho, T: t})
        # 2. cf: Coriolis, physics scheme
        state.update(tendency_step(self.ptis, state, self._coriolis, dt_s)[1])
        # 3. hs: IsentropicHorizontalSmoothing (moist), horizontal_smoothing.py:L172-L180
        sx, sy, sz = self.shape
        nb = max(self.smooth_order, hb.nb)
        for n in (S, SU, SV, QV, QC, QR):
            out = self.z()
            dwarfs.smoothing(self.smooth_order, state[n], self.gamma, out, (nb, nb, 0),
                             (sx - 2 * nb, sy - 2 * nb, sz))
            for o, d in (((0, 0, 0), (nb, sy, sz)), ((sx - nb, 0, 0), (nb, sy, sz)),
                         ((nb, 0, 0), (sx - 2 * nb, nb, sz)), ((nb, sy - nb, 0), (sx - 2 * nb, nb, sz))):
                dwarfs.copy(state[n], out, o, d)
            state[n] = out
        # 4. turb: IsentropicSmagorinsky, physics scheme
        state.update(tendency_step(self.ptis, state, self._smagorinsky, dt_s)[1])
        # 5. ivc: IsentropicVelocityComponents, isentropic/physics/diagnostics.py:L276-L301
        u, v = self.z(), self.z()
        dwarfs.get_velocity_components(nx, ny, nz, state[S], state[SU], state[SV], u, v)
        hb.set_outermost_layers_x(u, U)
        hb.set_outermost_layers_y(v, V)
        state.update({U: u, V: v})
        # 6. Kessler microphysics (+ t2d), 7. saturation adjustment (d2t, sa, t2d): physics scheme
        for fn in (self._kessler, self._saturation):
            diag, out = tendency_step(self.ptis, state, fn, dt_s)
            state.update(diag)
            state.update(out)
        # 8. vertical advection, rk3ws
        state.update(tendency_step("rk3ws", state, self._vertical_advection, dt_s)[1])
        # 9. sedimentation (rfv, sd), rk3ws
        diag, out = tendency_step("rk3ws", state, self._sedimentation, dt_s)
        state.update(diag)
        state.update(out)
        # 10. accumulated precipitation (rfv, ap): no tendencies, forward Euler wrapper
        diag, _ = tendency_step("forward_euler", state, lambda st: self._precipitation(st, dt_s), dt_s)
        state.update(diag)
        state["time"] = state["time"] + dt

    # ---------------------------------------------------------------- one model step
    def step(self, state, dt):
        """Loop body of driver_namelist_sus.py:L490-L512: topography, dynamics, carry-over of the
        fields the dycore does not return, physics.  Returns the new state dict."""
        self.nstep += 1
        self.topo.update(self.nstep * dt)
        out = self.dycore(state, {}, dt)
        new = {n: out[n].copy() for n in (S, SU, U, SV, V, QV, QC, QR)}
        for n in state:
            if n not in new and n != "time":
                new[n] = state[n]
        new["time"] = state["time"]
        self.physics(new, dt)
        return new
