# -*- coding: utf-8 -*-
"""Oracle (test infrastructure): Kessler microphysics, row K11 of SURVEY.md section 8a.

Follows the reference's numpy stencil definitions
  src/tasmania/physics/microphysics/kessler.py:L307-L376      kessler
  src/tasmania/physics/microphysics/kessler.py:L661-L714      saturation (diagnostic adjustment)
  src/tasmania/physics/microphysics/kessler.py:L981-L1032     saturation (prognostic)
  src/tasmania/physics/microphysics/kessler.py:L1183-L1203    fall_velocity
  src/tasmania/physics/microphysics/kessler.py:L1339-L1370    sedimentation
  src/tasmania/physics/microphysics/sedimentation_fluxes/first_order.py:L36-L43
  src/tasmania/physics/microphysics/sedimentation_fluxes/second_order.py:L36-L59
  src/tasmania/physics/microphysics/utils.py:L283-L305        accumulated_precipitation
  src/tasmania/framework/subclasses/subroutine_definitions/generics.py:L38-L40   set_output
Pinned bit for bit against the outputs of those definitions executed in place
(tests/golden/kessler.npz, tests/test_oracle_golden.py).
"""
import numpy as np

# default physical constants, kessler.py:L78-L82, L480-L486; utils.py:L147-L149
CONSTANTS = {"rd": 287.05, "rv": 461.52, "cp": 1004.0, "lhvw": 2.5e6, "rhow": 1.0e3}


def _box(origin, domain):
    i = slice(origin[0], origin[0] + domain[0])
    j = slice(origin[1], origin[1] + domain[1])
    k = slice(origin[2], origin[2] + domain[2])
    kp1 = slice(origin[2] + 1, origin[2] + domain[2] + 1)
    return i, j, k, kp1


def set_output(lhs, rhs, overwrite):
    lhs[...] = rhs if overwrite else lhs + rhs


def _p_exn(in_p, in_exn, i, j, k, kp1, on_interface_levels):
    if on_interface_levels:
        return 0.5 * (in_p[i, j, k] + in_p[i, j, kp1]), 0.5 * (in_exn[i, j, k] + in_exn[i, j, kp1])
    return in_p[i, j, k], in_exn[i, j, k]


def _qvs(t, p, beta):
    # Tetens' formula, kessler.py:L345-L348
    ps = 610.78 * np.exp(17.27 * (t - 273.16) / (t - 35.86))
    return beta * ps / p


def kessler(in_rho, in_p, in_t, in_exn, in_qc, in_qr, in_qv, out_qc_tnd, out_qr_tnd, out_qv_tnd,
            out_theta_tnd, *, a, k1, k2, ow_out_qc_tnd, ow_out_qr_tnd, ow_out_qv_tnd=True,
            ow_out_theta_tnd=True, origin, domain, air_pressure_on_interface_levels=True,
            rain_evaporation=True, beta=CONSTANTS["rd"] / CONSTANTS["rv"], lhvw=CONSTANTS["lhvw"]):
    i, j, k, kp1 = _box(origin, domain)
    p, exn = _p_exn(in_p, in_exn, i, j, k, kp1, air_pressure_on_interface_levels)
    qvs = _qvs(in_t[i, j, k], p, beta)
    ar = k1 * np.where(in_qc[i, j, k] > a, in_qc[i, j, k] - a, 0.0)
    with np.errstate(invalid="ignore"):
        cr = k2 * in_qc[i, j, k] * np.where(in_qr[i, j, k] > 0.0, in_qr[i, j, k] ** 0.875, 0.0)
        if rain_evaporation:
            er = np.where(
                in_qr[i, j, k] > 0.0,
                0.0484794 * (qvs - in_qv[i, j, k]) * (in_rho[i, j, k] * in_qr[i, j, k]) ** (13.0 / 20.0),
                0.0,
            )
    if not rain_evaporation:
        set_output(out_qc_tnd[i, j, k], -(ar + cr), ow_out_qc_tnd)
        set_output(out_qr_tnd[i, j, k], ar + cr, ow_out_qr_tnd)
    else:
        set_output(out_qv_tnd[i, j, k], er, ow_out_qv_tnd)
        set_output(out_qc_tnd[i, j, k], -(ar + cr), ow_out_qc_tnd)
        set_output(out_qr_tnd[i, j, k], ar + cr - er, ow_out_qr_tnd)
        set_output(out_theta_tnd[i, j, k], -lhvw / exn * er, ow_out_theta_tnd)


def _saturation_dq(in_p, in_t, in_exn, in_qv, in_qc, i, j, k, kp1, apoil, beta, lhvw, cp, rv):
    p, exn = _p_exn(in_p, in_exn, i, j, k, kp1, apoil)
    qvs = _qvs(in_t[i, j, k], p, beta)
    sat = (qvs - in_qv[i, j, k]) / (1.0 + qvs * (lhvw**2.0) / (cp * rv * (in_t[i, j, k] ** 2.0)))
    dq = np.where(sat <= in_qc[i, j, k], sat, in_qc[i, j, k])
    return dq, exn


def saturation_diagnostic(in_p, in_t, in_exn, in_qv, in_qc, out_qv, out_qc, out_t, tnd_theta, *, dt,
                          ow_tnd_theta, origin, domain, air_pressure_on_interface_levels=True,
                          beta=CONSTANTS["rd"] / CONSTANTS["rv"], lhvw=CONSTANTS["lhvw"],
                          cp=CONSTANTS["cp"], rv=CONSTANTS["rv"]):
    i, j, k, kp1 = _box(origin, domain)
    dq, exn = _saturation_dq(in_p, in_t, in_exn, in_qv, in_qc, i, j, k, kp1,
                             air_pressure_on_interface_levels, beta, lhvw, cp, rv)
    out_qv[i, j, k] = in_qv[i, j, k] + dq
    out_qc[i, j, k] = in_qc[i, j, k] - dq
    out_t[i, j, k] = in_t[i, j, k] - dq * lhvw / cp
    set_output(tnd_theta[i, j, k], (lhvw / exn) * (-dq / dt), ow_tnd_theta)


def saturation_prognostic(in_p, in_t, in_exn, in_qv, in_qc, tnd_qv, tnd_qc, tnd_theta, *, sr, origin,
                          domain, ow_tnd_qv, ow_tnd_qc, ow_tnd_theta,
                          air_pressure_on_interface_levels=True,
                          beta=CONSTANTS["rd"] / CONSTANTS["rv"], lhvw=CONSTANTS["lhvw"],
                          cp=CONSTANTS["cp"], rv=CONSTANTS["rv"]):
    i, j, k, kp1 = _box(origin, domain)
    dq, exn = _saturation_dq(in_p, in_t, in_exn, in_qv, in_qc, i, j, k, kp1,
                             air_pressure_on_interface_levels, beta, lhvw, cp, rv)
    set_output(tnd_qv[i, j, k], sr * dq, ow_tnd_qv)
    set_output(tnd_qc[i, j, k], -sr * dq, ow_tnd_qc)
    set_output(tnd_theta[i, j, k], -sr * (lhvw / exn) * dq, ow_tnd_theta)


def fall_velocity(in_rho, in_rho_s, in_qr, out_vt, *, origin, domain):
    i, j, k, _ = _box(origin, domain)
    out_vt[i, j, k] = (
        36.34
        * (1.0e-3 * in_rho[i, j, k] * np.where(in_qr[i, j, k] > 0.0, in_qr[i, j, k], 0.0)) ** 0.1346
        * (in_rho_s[i, j, k] / in_rho[i, j, k]) ** 0.5
    )


def sedimentation_flux(order, rho, h, q, vt):
    """d(rho q vt)/dz by first / second order upwind differences over the variable height h."""
    if order == 1:
        return (
            rho[:, :, :-1] * q[:, :, :-1] * vt[:, :, :-1] - rho[:, :, 1:] * q[:, :, 1:] * vt[:, :, 1:]
        ) / (h[:, :, :-1] - h[:, :, 1:])
    tmp_a = (2.0 * h[:, :, 2:] - h[:, :, 1:-1] - h[:, :, :-2]) / (
        (h[:, :, 1:-1] - h[:, :, 2:]) * (h[:, :, :-2] - h[:, :, 2:])
    )
    tmp_b = (h[:, :, :-2] - h[:, :, 2:]) / (
        (h[:, :, 1:-1] - h[:, :, 2:]) * (h[:, :, :-2] - h[:, :, 1:-1])
    )
    tmp_c = (h[:, :, 2:] - h[:, :, 1:-1]) / (
        (h[:, :, :-2] - h[:, :, 2:]) * (h[:, :, :-2] - h[:, :, 1:-1])
    )
    return (
        tmp_a * rho[:, :, 2:] * q[:, :, 2:] * vt[:, :, 2:]
        + tmp_b * rho[:, :, 1:-1] * q[:, :, 1:-1] * vt[:, :, 1:-1]
        + tmp_c * rho[:, :, :-2] * q[:, :, :-2] * vt[:, :, :-2]
    )


def sedimentation(in_rho, in_h, in_qr, in_vt, out_tnd_qr, *, ow_out_tnd_qr, origin, domain, order):
    i = slice(origin[0], origin[0] + domain[0])
    j = slice(origin[1], origin[1] + domain[1])
    kb, ke = origin[2], origin[2] + domain[2]
    ext = order  # sflux.nb
    h = 0.5 * (in_h[i, j, kb:ke] + in_h[i, j, kb + 1 : ke + 1])
    dfdz = sedimentation_flux(order, in_rho[i, j, kb:ke], h, in_qr[i, j, kb:ke], in_vt[i, j, kb:ke])
    if ow_out_tnd_qr:
        out_tnd_qr[i, j, kb : kb + ext] = 0.0
    set_output(out_tnd_qr[i, j, kb + ext : ke], dfdz / in_rho[i, j, kb + ext : ke], ow_out_tnd_qr)


def accumulated_precipitation(in_rho, in_qr, in_vt, in_accprec, out_prec, out_accprec, *, dt, origin,
                              domain, rhow=CONSTANTS["rhow"]):
    i, j, k, _ = _box(origin, domain)
    out_prec[i, j, k] = 3.6e6 * in_rho[i, j, k] * in_qr[i, j, k] * in_vt[i, j, k] / rhow
    out_accprec[i, j, k] = in_accprec[i, j, k] + dt * out_prec[i, j, k] / 3.6e3
