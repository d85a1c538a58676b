# -*- coding: utf-8 -*-
"""tb200_velocity_components (both staggered velocity components + outermost faces in one pass)
against the oracle's velocity_x / velocity_y (dwarfs/diagnostics.py:L219-L272) and the Relaxed
boundary's outermost layers (relaxed.py:L161-L191): bit for bit, odd and even row lengths, strips
of 64 and of 8 rows, zero and tiny momenta (the fast path of the kernels' own division)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nx,ny,nz", [(2, 2, 1), (7, 5, 3), (64, 9, 2), (65, 70, 4), (130, 131, 5), (161, 161, 60),
                                      (600, 140, 64)])
@pytest.mark.parametrize("with_ref", [True, False])
def test_velocity_components_vs_oracle(nx, ny, nz, with_ref):
    import tasmania_b200 as tb
    from oracle import dwarfs
    from tasmania_b200 import lib

    rng = np.random.default_rng(nx * 1000 + ny)
    shape = (nx + 1, ny + 1, nz + 1)
    s = rng.uniform(10.0, 1000.0, shape)
    su = s * rng.uniform(-50.0, 50.0, shape)
    sv = s * rng.uniform(-50.0, 50.0, shape)
    sv[: nx // 2] = 0.0                      # v = 0 upstream: zero numerators
    sv[nx // 2: nx // 2 + 1] *= 1e-250      # ... and a front of tiny ones
    u0, v0 = rng.standard_normal(shape), rng.standard_normal(shape)
    ur, vr = rng.standard_normal(shape), rng.standard_normal(shape)
    want_u, want_v = u0.copy(), v0.copy()
    dwarfs.get_velocity_components(nx, ny, nz, s, su, sv, want_u, want_v)
    if with_ref:
        want_u[0, :ny, :nz], want_u[nx, :ny, :nz] = ur[0, :ny, :nz], ur[nx, :ny, :nz]
        want_v[:nx, 0, :nz], want_v[:nx, ny, :nz] = vr[:nx, 0, :nz], vr[:nx, ny, :nz]
    f = lib.as_field
    d = {n: tb.as_storage(a) for n, a in (("s", s), ("su", su), ("sv", sv), ("u", u0), ("v", v0), ("ur", ur), ("vr", vr))}
    rc = lib.load().tb200_velocity_components(f(d["s"]), f(d["su"]), f(d["sv"]), f(d["u"]), f(d["v"]),
                                              f(d["ur"]) if with_ref else None, f(d["vr"]) if with_ref else None,
                                              nx, ny, nz, lib.current_stream())
    lib.check(rc, "tb200_velocity_components")
    np.testing.assert_array_equal(tb.to_numpy(d["u"]), want_u)
    np.testing.assert_array_equal(tb.to_numpy(d["v"]), want_v)
