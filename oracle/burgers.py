# -*- coding: utf-8 -*-
"""Oracle (test infrastructure): the 2-D Burgers dwarf, row K10 of SURVEY.md section 8a.

Follows
  src/tasmania/burgers/dynamics/subclasses/advection/{first..sixth}_order.py  (advection)
  src/tasmania/burgers/dynamics/stepper.py:L188-L227                          (forward_euler)
  src/tasmania/burgers/dynamics/subclasses/stepper/{forward_euler,rk2,rk3ws}.py
  src/tasmania/burgers/dynamics/dycore.py:L158-L173                           (stage + boundary)
  src/tasmania/burgers/state.py:L59-L152                                      (Zhao solution)
"""

import numpy as np

EXTENT = {1: 1, 2: 1, 3: 2, 4: 2, 5: 3, 6: 3}
ORDER = {"first_order": 1, "second_order": 2, "third_order": 3, "fourth_order": 4,
         "fifth_order": 5, "sixth_order": 6}


def _shift(a, e, d, axis):
    """a[e+d : -e+d] along ``axis`` and a[e:-e] along the other horizontal axis."""
    n = a.shape[axis]
    idx = [slice(e, a.shape[0] - e), slice(e, a.shape[1] - e)]
    idx[axis] = slice(e + d, n - e + d)
    return a[tuple(idx)]


def _term(order, a, q, axis, dd):
    """One advective term ``a * dq/dx`` of the given order on the interior of ``q``."""
    e = EXTENT[order]
    c = _shift(a, e, 0, axis)
    s = lambda d: _shift(q, e, d, axis)  # noqa: E731
    if order == 1:
        return c / (2.0 * dd) * (s(1) - s(-1)) - np.abs(c) / (2.0 * dd) * (s(1) - 2.0 * s(0) + s(-1))
    if order == 2:
        return c / (2.0 * dd) * (s(1) - s(-1))
    if order == 3:
        return c / (12.0 * dd) * (8.0 * (s(1) - s(-1)) - (s(2) - s(-2))) + np.abs(c) / (12.0 * dd) * (
            s(2) + s(-2) - 4.0 * (s(1) + s(-1)) + 6.0 * s(0)
        )
    if order == 4:
        return c / (12.0 * dd) * (8.0 * (s(1) - s(-1)) - (s(2) - s(-2)))
    if order == 5:
        return c / (60.0 * dd) * (
            +45.0 * (s(1) - s(-1)) - 9.0 * (s(2) - s(-2)) + (s(3) - s(-3))
        ) - np.abs(c) / (60.0 * dd) * (
            +(s(3) + s(-3)) - 6.0 * (s(2) + s(-2)) + 15.0 * (s(1) + s(-1)) - 20.0 * s(0)
        )
    if order == 6:
        return c / (60.0 * dd) * (+45.0 * (s(1) - s(-1)) - 9.0 * (s(2) - s(-2)) + (s(3) - s(-3)))
    raise ValueError(order)


def advection(order, dx, dy, u, v):
    """(adv_u_x, adv_u_y, adv_v_x, adv_v_y) on the interior of the input box."""
    return (_term(order, u, u, 0, dx), _term(order, v, u, 1, dy),
            _term(order, u, v, 0, dx), _term(order, v, v, 1, dy))


def forward_euler(order, in_u, in_v, in_u_tmp, in_v_tmp, out_u, out_v, *, dt, dx, dy, origin,
                  domain, u_tnd=None, v_tnd=None):
    """stepper.py:L188-L227."""
    e = EXTENT[order]
    i0, i1 = origin[0], origin[0] + domain[0]
    j0, j1 = origin[1], origin[1] + domain[1]
    k = slice(origin[2], origin[2] + domain[2])
    i, j = slice(i0, i1), slice(j0, j1)
    iext, jext = slice(i0 - e, i1 + e), slice(j0 - e, j1 + e)
    aux, auy, avx, avy = advection(order, dx, dy, in_u_tmp[iext, jext, k], in_v_tmp[iext, jext, k])
    if u_tnd is not None:
        out_u[i, j, k] = in_u[i, j, k] - dt * (aux + auy - u_tnd[i, j, k])
    else:
        out_u[i, j, k] = in_u[i, j, k] - dt * (aux + auy)
    if v_tnd is not None:
        out_v[i, j, k] = in_v[i, j, k] - dt * (avx + avy - v_tnd[i, j, k])
    else:
        out_v[i, j, k] = in_v[i, j, k] - dt * (avx + avy)


def zhao_solution(t, x, y, eps, field_name, nz=1):
    """state.py:L97-L152; x, y 1-D coordinate arrays [m], t seconds since the initial time."""
    x = np.tile(x[:, None, None], (1, len(y), nz))
    y = np.tile(y[None, :, None], (x.shape[0], 1, nz))
    if field_name == "x_velocity":
        return (
            -2.0 * eps * 2.0 * np.pi * np.exp(-5.0 * np.pi**2 * eps * t)
            * np.cos(2.0 * np.pi * x) * np.sin(np.pi * y)
            / (2.0 + np.exp(-5.0 * np.pi**2 * eps * t) * np.sin(2.0 * np.pi * x) * np.sin(np.pi * y))
        )
    if field_name == "y_velocity":
        return (
            -2.0 * eps * np.pi * np.exp(-5.0 * np.pi**2 * eps * t)
            * np.sin(2.0 * np.pi * x) * np.cos(np.pi * y)
            / (2.0 + np.exp(-5.0 * np.pi**2 * eps * t) * np.sin(2.0 * np.pi * x) * np.sin(np.pi * y))
        )
    raise ValueError(field_name)


class BurgersDycore:
    """Raw-array restatement of ``BurgersDynamicalCore`` with a Dirichlet (Zhao) or any
    oracle.boundary lateral boundary.  ``dirichlet``: callable (elapsed_seconds, sx, sy, name)
    -> rim values, or None to use ``hb.enforce_raw``."""

    STAGES = {"forward_euler": 1, "rk2": 2, "rk3ws": 3}

    def __init__(self, nx, ny, dx, dy, nb, scheme="rk3ws", flux="third_order", hb=None,
                 dirichlet=None):
        self.nx, self.ny, self.dx, self.dy, self.nb = nx, ny, dx, dy, nb
        self.scheme, self.order = scheme, ORDER[flux]
        assert nb >= EXTENT[self.order]
        self.stages = self.STAGES[scheme]
        self.hb, self.dirichlet = hb, dirichlet
        self._outs = None

    def _dts(self, stage, timestep):
        """(time-label increment, dt [s]) -- rk3ws.py:L48-L58, rk2.py:L44-L50."""
        ts = timestep.total_seconds()
        if self.scheme == "forward_euler":
            return timestep, ts
        if self.scheme == "rk2":
            return 0.5 * timestep, (0.5 * ts if stage == 0 else ts)
        if stage == 0:
            return 1.0 / 3.0 * timestep, 1.0 / 3.0 * ts
        if stage == 1:
            return 1.0 / 6.0 * timestep, 0.5 * ts
        return 1.0 / 2.0 * timestep, ts

    def _enforce(self, out):
        nb = self.nb
        if self.dirichlet is not None:
            for name in ("x_velocity", "y_velocity"):
                f = out[name]
                mi, mj = self.nx, self.ny
                for sx, sy in ((slice(0, nb), slice(0, mj)), (slice(mi - nb, mi), slice(0, mj)),
                               (slice(nb, mi - nb), slice(0, nb)),
                               (slice(nb, mi - nb), slice(mj - nb, mj))):
                    f[sx, sy] = self.dirichlet(out["time"], sx, sy, name)
        elif self.hb is not None:
            self.hb.enforce_raw(out, ("x_velocity", "y_velocity"))

    def __call__(self, state, tendencies, timestep):
        shape = state["x_velocity"].shape
        if self._outs is None:
            self._outs = [{n: np.zeros(shape) for n in ("x_velocity", "y_velocity")}
                          for _ in range(self.stages)]
        nb, nx, ny = self.nb, self.nx, self.ny
        cur = state
        for stage in range(self.stages):
            out = self._outs[stage] if stage < self.stages - 1 else {
                n: np.zeros(shape) for n in ("x_velocity", "y_velocity")}
            dtr, dt = self._dts(stage, timestep)
            forward_euler(self.order, state["x_velocity"], state["y_velocity"], cur["x_velocity"],
                          cur["y_velocity"], out["x_velocity"], out["y_velocity"], dt=dt,
                          dx=self.dx, dy=self.dy, origin=(nb, nb, 0),
                          domain=(nx - 2 * nb, ny - 2 * nb, 1),
                          u_tnd=(tendencies or {}).get("x_velocity"),
                          v_tnd=(tendencies or {}).get("y_velocity"))
            out["time"] = cur["time"] + dtr
            self._enforce(out)
            cur = out
        cur["time"] = state["time"] + timestep
        return cur
