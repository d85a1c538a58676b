# -*- coding: utf-8 -*-
"""BASELINE configs[2] on the GPU: the moist isentropic model (dynamical core + Kessler
microphysics with sedimentation under sequential-update splitting, tasmania_b200.isentropic_moist)
against the oracle's straight-line restatement (oracle/moist_model.py) on the same seeded state,
and the coupler-glue kernel ``tb200_fma_fields`` against numpy (bit-exact).

Tolerance of the model run: 1e-12 relative (max-norm per field) -- north_star's bound -- on every
prognostic and diagnostic field; 1e-11 on the promoted latent-heating rate
(tendency_of_air_potential_temperature), which is a difference of nearly equal numbers
(qvs - qv at 98 % relative humidity) and so amplifies the last-ulp differences between CUDA's and
glibc's exp / pow.  Measured on B200: <= 1.5e-14 on the fields, <= 1.6e-12 on the heating rate.
"""
from datetime import datetime, timedelta

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-12
RTOL_HEATING = 1e-11


def test_fma_fields_bitwise():
    import tasmania_b200 as tb
    from tasmania_b200 import stencils

    rng = np.random.default_rng(11)
    shape = (37, 21, 9)
    nf = 11  # more than TB200_FMA_MAX_FIELDS: two launches
    a = [rng.standard_normal(shape) for _ in range(nf)]
    b = [rng.standard_normal(shape) for _ in range(nf)]
    f = 5.0 / 3.0
    da, db = [tb.as_storage(x) for x in a], [tb.as_storage(x) for x in b]
    out = [tb.zeros(shape) for _ in range(nf)]
    stencils.fma_fields(out, da, db, f, origin=(0, 0, 0), domain=shape)
    for n in range(nf):
        np.testing.assert_array_equal(tb.to_numpy(out[n]), a[n] + f * b[n])
    # sub-box, in place on a
    stencils.fma_fields(da[:3], da[:3], db[:3], -0.25, origin=(2, 3, 1), domain=(30, 10, 5))
    for n in range(3):
        want = a[n].copy()
        want[2:32, 3:13, 1:6] = a[n][2:32, 3:13, 1:6] + -0.25 * b[n][2:32, 3:13, 1:6]
        np.testing.assert_array_equal(tb.to_numpy(da[n]), want)
    with pytest.raises(tb.lib.B200Error):
        stencils.fma_fields(out[:1], da[:1], db[:1], 1.0, origin=(0, 0, 0), domain=(38, 21, 9))


@pytest.mark.parametrize("dims,nsteps", [((33, 29, 14), 12), ((49, 41, 20), 6)])
def test_moist_model_vs_oracle(dims, nsteps):
    import tasmania_b200 as tb
    from oracle import boundary as ob
    from oracle import isentropic as oi
    from oracle import moist_model as mm
    from tasmania_b200.isentropic_moist import IsentropicMoistSUS
    from tests import helpers as hp

    nx, ny, nz = dims
    dt = timedelta(seconds=5)
    grid, np_state = hp.moist_case(nx, ny, nz)
    # ---- oracle
    ogrid = oi.Grid(nx, ny, nz, grid.dx, grid.dy, grid.dz, grid.z_on_interface_levels, grid.z)
    ohb = ob.Relaxed(nx, ny, nz, 3, 6)
    ohb.reference_state = {n: v.copy() for n, v in np_state.items()}
    otopo = hp.Topography(grid.topography.steady_profile, 60.0)
    pt = float(np_state[mm.P][0, 0, 0])
    omodel = mm.MoistIsentropicModel(ogrid, ohb, otopo, pt, damp_depth=4)
    ost = {n: v.copy() for n, v in np_state.items()}
    ost[mm.W] = np.zeros_like(ost[mm.S])
    ost["time"] = datetime(1992, 2, 20)
    # ---- GPU
    model = IsentropicMoistSUS(grid, np_state, dt, damp_depth=4)
    launches0 = tb.lib.launch_count()
    for _ in range(nsteps):
        ost = omodel.step(ost, dt)
        model.step()
    assert tb.lib.launch_count() - launches0 > 50 * nsteps  # the CUDA library did the work
    final = model.state
    assert final["time"] == ost["time"]
    worst = {}
    for n in (mm.S, mm.SU, mm.SV, mm.U, mm.V, mm.MTG, mm.P, mm.EXN, mm.H, mm.RHO, mm.T, mm.QV,
              mm.QC, mm.QR, mm.W, mm.VT):
        worst[n] = hp.relerr(tb.to_numpy(final[n])[:nx, :ny, :nz], ost[n][:nx, :ny, :nz])
    for n in (mm.PREC, mm.ACCPREC):
        worst[n] = hp.relerr(tb.to_numpy(final[n])[:nx, :ny], ost[n][:nx, :ny])
    print(f"moist model {dims}, {nsteps} steps, relative errors:",
          {k: float(f"{v:.2e}") for k, v in worst.items()})
    assert worst.pop(mm.W) <= RTOL_HEATING
    assert max(worst.values()) <= RTOL, worst
    # all the moist processes were active in the compared run
    box = (slice(0, nx), slice(0, ny), slice(0, nz))
    assert float(ost[mm.QC][box].max()) > 1e-5 and float(ost[mm.QR][box].max()) > 1e-5
    assert float(ost[mm.ACCPREC].max()) > 0.0 and float(np.abs(ost[mm.W][box]).max()) > 0.0


@pytest.mark.parametrize("scheme", ["upwind", "centered", "third_order_upwind", "fifth_order_upwind"])
@pytest.mark.parametrize("moist", [False, True])
def test_vertical_advection_step_is_advection_plus_fma_bitwise(scheme, moist):
    """``tb200_vertical_advection_step`` (tendency kernel + stage update in one) against the
    tendencies of ``tb200_vertical_advection`` followed by ``base + factor * tendency`` in numpy."""
    import tasmania_b200 as tb
    from tasmania_b200.grid import Grid
    from tasmania_b200.isentropic_physics import IsentropicVerticalAdvection, W_ML

    rng = np.random.default_rng(21)
    nx, ny, nz = 19, 11, 16
    grid = Grid((0.0, 1.0), nx, (0.0, 1.0), ny, (340.0, 280.0), nz)
    shape = (nx + 1, ny + 1, nz + 1)
    comp = IsentropicVerticalAdvection(grid, flux_scheme=scheme, moist=moist)
    names = comp.tendency_names
    host = {n: rng.uniform(0.5, 2.0, shape) for n in names}
    host[W_ML] = rng.standard_normal(shape) * 1e-2
    base = {n: rng.standard_normal(shape) for n in names}
    state = {n: tb.as_storage(v) for n, v in host.items()}
    dbase = {n: tb.as_storage(v) for n, v in base.items()}
    tnd = {n: tb.as_storage(rng.standard_normal(shape)) for n in names}  # garbage: overwritten
    comp.array_call(state, tnd, {}, {n: True for n in names})
    out = {n: tb.as_storage(rng.standard_normal(shape)) for n in names}
    factor = 5.0 / 3.0
    comp.array_call_stepped(state, dbase, factor, out)
    for n in names:
        np.testing.assert_array_equal(tb.to_numpy(out[n]), base[n] + factor * tb.to_numpy(tnd[n]),
                                      err_msg=n)
    with pytest.raises(tb.lib.B200Error):  # outputs must not alias the inputs
        comp.array_call_stepped(state, dbase, factor, state)


def test_fused_paths_equal_the_reference_shaped_paths_bitwise(monkeypatch):
    """Moist model with the b200 fusions (fused moist dycore stage, stage update inside the vertical
    advection kernel, frame relaxation of all fields in one launch) against the same model issuing the reference's call
    sequence (tendencies + fma, one full-box irelax per field): identical bits."""
    import tasmania_b200 as tb
    from tasmania_b200.isentropic_moist import IsentropicMoistSUS
    from tests import helpers as hp

    nx, ny, nz, nsteps = 33, 29, 14, 5
    res = []
    for fused in (True, False):
        monkeypatch.setenv("TB200_FUSED_STEP", "1" if fused else "0")
        monkeypatch.setenv("TB200_RELAX", "frame" if fused else "full")
        monkeypatch.setenv("TB200_MOIST_FUSED", "1" if fused else "0")  # the fused moist dycore stage
        grid, np_state = hp.moist_case(nx, ny, nz)
        model = IsentropicMoistSUS(grid, np_state, timedelta(seconds=5), damp_depth=4)
        n0 = tb.lib.launch_count()
        model.run(nsteps)
        res.append(({n: tb.to_numpy(v) for n, v in model.state.items() if n != "time"},
                    (tb.lib.launch_count() - n0) / nsteps))
    assert res[0][1] < res[1][1] - 40  # 63 against 122 launches per step
    for n, v in res[1][0].items():
        np.testing.assert_array_equal(res[0][0][n], v, err_msg=n)


def test_moist_model_vs_reference_fixture():
    """The GPU model against tests/golden/moist_model.npz: five steps of the moist benchmark loop
    computed by the REFERENCE ITSELF (its moist dynamical core and physics suite run in place, numpy
    backend; tests/golden/generate_moist_model.py).  Same tolerances as against the oracle, which
    reproduces this fixture bit for bit (tests/test_moist_model_oracle.py)."""
    import tasmania_b200 as tb
    from oracle import moist_model as mm
    from tasmania_b200.isentropic_moist import IsentropicMoistSUS
    from tests import helpers as hp

    fx = hp.load("moist_model")
    nx, ny, nz, nb, nr, nsteps, damp_depth = (int(v) for v in fx["dims"])
    dt_s, max_height, topo_seconds, rh = (float(v) for v in fx["params"])
    grid, np_state = hp.moist_case(nx, ny, nz, max_height=max_height, topo_seconds=topo_seconds,
                                   relative_humidity=rh)
    # start from the fixture's own initial state (host libm may differ in the last bit between
    # the machine that wrote the fixture and this one; the case builder only supplies the grid)
    np_state = {n: fx["init_" + n] for n in np_state}
    model = IsentropicMoistSUS(grid, np_state, timedelta(seconds=dt_s), nb=nb, nr=nr,
                               damp_depth=damp_depth)
    final = model.run(nsteps)
    worst = {}
    for key in fx.files:
        if not key.startswith("final_"):
            continue
        n = key[6:]
        ref = fx[key]
        box = (slice(0, nx), slice(0, ny), slice(0, nz if ref.shape[2] > 1 else 1))
        worst[n] = hp.relerr(tb.to_numpy(final[n])[box], ref[box])
    print("moist model vs the reference's own run, relative errors:",
          {k: float(f"{v:.2e}") for k, v in worst.items()})
    assert len(worst) >= 18
    assert worst.pop(mm.W) <= RTOL_HEATING
    assert max(worst.values()) <= RTOL, worst
