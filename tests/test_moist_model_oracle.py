# -*- coding: utf-8 -*-
"""The oracle's moist model (oracle/moist_model.py: dynamical core + sequential-update-splitting
physics of BASELINE config 3) on the CPU: the pieces it is built of are pinned bit for bit on the
reference (tests/test_oracle_golden.py); here the coupling logic is checked through properties
the reference's couplers have by construction."""
from datetime import datetime, timedelta

import numpy as np

from oracle import boundary as ob
from oracle import isentropic as oi
from oracle import moist_model as mm
from tests import helpers as hp


def build(nx=21, ny=19, nz=10, **kw):
    grid, state = hp.moist_case(nx, ny, nz, **kw)
    ogrid = oi.Grid(nx, ny, nz, grid.dx, grid.dy, grid.dz, grid.z_on_interface_levels, grid.z)
    hb = ob.Relaxed(nx, ny, nz, 3, 6)
    hb.reference_state = {n: v.copy() for n, v in state.items()}
    topo = hp.Topography(grid.topography.steady_profile, 60.0)
    pt = float(state[mm.P][0, 0, 0])
    model = mm.MoistIsentropicModel(ogrid, hb, topo, pt, damp_depth=4)
    st = {n: v.copy() for n, v in state.items()}
    st[mm.W] = np.zeros_like(st[mm.S])
    st["time"] = datetime(1992, 2, 20)
    return model, st


def test_tendency_step_is_the_textbook_scheme():
    """dy/dt = -y: forward Euler, midpoint RK2 and the Wicker-Skamarock RK3 reduce to their
    stability polynomials 1 - z, 1 - z + z^2/2, 1 - z + z^2/2 - z^3/6."""
    y0, dt = np.full((2, 2, 2), 3.0), 0.1
    fn = lambda st: ({"y": -st["y"]}, {"seen": st["y"].copy()})  # noqa: E731
    for scheme, poly in (("forward_euler", 1 - dt), ("rk2", 1 - dt + dt**2 / 2),
                         ("rk3ws", 1 - dt + dt**2 / 2 - dt**3 / 6)):
        diag, out = mm.tendency_step(scheme, {"y": y0, "other": y0}, fn, dt)
        np.testing.assert_allclose(out["y"], y0 * poly, rtol=1e-15)
        assert set(out) == {"y"}
        np.testing.assert_array_equal(diag["seen"], y0)  # diagnostics of the FIRST stage


def test_moist_model_runs_and_all_processes_are_active():
    model, st = build()
    dt = timedelta(seconds=5)
    acc0 = st[mm.ACCPREC].copy()
    qtot0 = sum(float((st[mm.S] * st[q])[:21, :19, :10].sum()) for q in (mm.QV, mm.QC, mm.QR))
    for _ in range(6):
        st = model.step(st, dt)
    for n, v in st.items():
        if n != "time":
            assert np.isfinite(v).all(), n
    assert st["time"] == datetime(1992, 2, 20) + 6 * dt
    nx, ny, nz = 21, 19, 10
    box = (slice(0, nx), slice(0, ny), slice(0, nz))
    # clipping keeps the mass fractions non-negative up to the physics' own increments
    assert float(st[mm.QC][box].max()) > 1e-5 and float(st[mm.QR][box].max()) > 1e-5
    assert float(np.abs(st[mm.W][box]).max()) > 0.0        # latent heating promoted to the state
    assert float(st[mm.VT][box].max()) > 1.0               # rain falls at metres per second
    assert float((st[mm.ACCPREC] - acc0).max()) >= 0.0
    assert float(st[mm.PREC].max()) >= 0.0
    assert float(np.abs(st[mm.SV][box]).max()) > 0.0       # Coriolis / mountain turned the flow
    qtot = sum(float((st[mm.S] * st[q])[box].sum()) for q in (mm.QV, mm.QC, mm.QR))
    assert abs(qtot - qtot0) / qtot0 < 0.05                # water is moved around, not created


def test_physics_leaves_untouched_what_no_component_outputs():
    """SequentialUpdateSplitting only replaces the outputs of its components."""
    model, st = build()
    dt = timedelta(seconds=5)
    st2 = dict(st)
    st2["bystander"] = np.arange(8.0).reshape(2, 2, 2)
    model.physics(st2, dt)
    np.testing.assert_array_equal(st2["bystander"], np.arange(8.0).reshape(2, 2, 2))
    # u, v are re-diagnosed from the smoothed momenta inside the physics
    assert not np.array_equal(st2[mm.U], st[mm.U])
    # zero timestep: tendency components do nothing, diagnostics and smoothing still act
    model2, st3 = build()
    before = {n: st3[n].copy() for n in (mm.QV, mm.QC, mm.QR)}
    model2.ptis = "rk2"
    st4 = dict(st3)
    model2.physics(st4, timedelta(seconds=0))
    sm = {}
    for n in before:
        out = np.zeros_like(before[n])
        from oracle import dwarfs
        sx, sy, sz = before[n].shape
        dwarfs.smoothing(2, before[n], model2.gamma, out, (3, 3, 0), (sx - 6, sy - 6, sz))
        sm[n] = out
        np.testing.assert_array_equal(st4[n][3:-3, 3:-3], out[3:-3, 3:-3])


def test_oracle_moist_model_reproduces_the_reference_fixture():
    """tests/golden/moist_model.npz was produced by the reference itself (five steps of the moist
    benchmark loop, tests/golden/generate_moist_model.py); the oracle reproduces every field of it
    bit for bit from the fixture's initial state."""
    fx = hp.load("moist_model")
    nx, ny, nz, nb, nr, nsteps, damp_depth = (int(v) for v in fx["dims"])
    dt_s, max_height, topo_seconds, rh = (float(v) for v in fx["params"])
    model, st = build(nx, ny, nz, max_height=max_height, topo_seconds=topo_seconds,
                      relative_humidity=rh)
    names = [k[5:] for k in fx.files if k.startswith("init_")]
    for n in names:  # the case builder gives the fixture's initial state
        np.testing.assert_array_equal(st[n], fx["init_" + n], err_msg=n)
    for _ in range(nsteps):
        st = model.step(st, timedelta(seconds=dt_s))
    assert {k[6:] for k in fx.files if k.startswith("final_")} == set(st) - {"time"}
    for n in st:
        if n != "time":
            np.testing.assert_array_equal(st[n], fx["final_" + n], err_msg=n)
