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
