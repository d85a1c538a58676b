# -*- coding: utf-8 -*-
"""Oracle (test infrastructure): horizontal numerical fluxes of the isentropic core.

Restates the point formulas of
  src/tasmania/isentropic/dynamics/subclasses/horizontal_fluxes/upwind.py:L32-L39
  .../horizontal_fluxes/centered.py:L173-L202
  .../horizontal_fluxes/third_order_upwind.py:L32-L55
  .../horizontal_fluxes/fifth_order_upwind.py:L32-L75
as used by the "minimal" flux classes
  .../minimal_horizontal_fluxes/{upwind,centered,third_order_upwind,fifth_order_upwind}.py.

Convention: ``face_flux(w, phi, axis)`` returns an array shorter
than the input by ``2*extent`` along ``axis``; entry ``m`` is the flux through face
``m + extent``, i.e. ``w[f] * Phi(phi[f-e .. f+e-1])`` with ``f = m + e``.
"""
import numpy as np

EXTENT = {"upwind": 1, "centered": 1, "third_order_upwind": 2, "fifth_order_upwind": 3}
ORDER = {"upwind": 1, "centered": 2, "third_order_upwind": 3, "fifth_order_upwind": 5}


def _sl(arr, axis, lo, hi):
    """arr[lo:hi] along ``axis`` where hi <= 0 counts from the end (0 = to the end)."""
    idx = [slice(None)] * arr.ndim
    idx[axis] = slice(lo, hi if hi != 0 else None)
    return arr[tuple(idx)]


def face_flux(scheme, w, phi, axis):
    """Flux of ``phi`` advected by the staggered velocity ``w`` along ``axis`` (0: x, 1: y)."""
    if scheme == "upwind":
        # upwind.py:L32-L39 -- strict ``> 0`` picks the upstream cell
        wf = _sl(w, axis, 1, -1)
        return wf * np.where(wf > 0.0, _sl(phi, axis, 0, -2), _sl(phi, axis, 1, -1))
    if scheme == "centered":
        # centered.py:L173-L202 -- (w * 0.5) * (phi[f-1] + phi[f])
        wf = _sl(w, axis, 1, -1)
        return wf * 0.5 * (_sl(phi, axis, 0, -2) + _sl(phi, axis, 1, -1))
    if scheme == "third_order_upwind":
        # third_order_upwind.py:L32-L55
        wf = _sl(w, axis, 2, -2)
        p0, m1 = _sl(phi, axis, 2, -2), _sl(phi, axis, 1, -3)
        p1, m2 = _sl(phi, axis, 3, -1), _sl(phi, axis, 0, -4)
        flux4 = wf / 12.0 * (7.0 * (p0 + m1) - (p1 + m2))
        return flux4 - np.abs(wf) / 12.0 * (3.0 * (p0 - m1) - (p1 - m2))
    if scheme == "fifth_order_upwind":
        # fifth_order_upwind.py:L32-L75
        wf = _sl(w, axis, 3, -3)
        p0, m1 = _sl(phi, axis, 3, -3), _sl(phi, axis, 2, -4)
        p1, m2 = _sl(phi, axis, 4, -2), _sl(phi, axis, 1, -5)
        p2, m3 = _sl(phi, axis, 5, -1), _sl(phi, axis, 0, -6)
        flux6 = wf / 60.0 * (37.0 * (p0 + m1) - 8.0 * (p1 + m2) + (p2 + m3))
        return flux6 - np.abs(wf) / 60.0 * (10.0 * (p0 - m1) - 5.0 * (p1 - m2) + (p2 - m3))
    raise ValueError(f"unknown horizontal flux scheme {scheme!r}")
