# -*- coding: utf-8 -*-
"""The host side of the product run NUMERICALLY without a GPU: the library is replaced by
tests/abi_oracle.py:OracleStub, which decodes every C-ABI call like the C side does and carries it
out with the oracle on host buffers.  Mirrors, coupling layer, model assembly, marshalling of
strides / flags / boxes, buffer rotation and CUDA-graph bookkeeping are thereby compared with the
oracle's own models bit for bit (both sides are numpy).  The CUDA code itself is what the
``-m gpu`` tests hold to the same oracle."""
from datetime import datetime, timedelta

import numpy as np
import pytest

from oracle import boundary as ob
from oracle import isentropic as oi
from oracle import moist_model as mm
from tests import helpers as hp
from tests.abi_oracle import OracleStub
from tests.abi_stub import FakeCapture, stubbed_library


def _oracle_moist(nx, ny, nz, grid, np_state):
    ogrid = oi.Grid(nx, ny, nz, grid.dx, grid.dy, grid.dz, grid.z_on_interface_levels, grid.z)
    ohb = ob.Relaxed(nx, ny, nz, 3, 6)
    ohb.reference_state = {n: v.copy() for n, v in np_state.items()}
    otopo = hp.Topography(grid.topography.steady_profile, 60.0)
    model = mm.MoistIsentropicModel(ogrid, ohb, otopo, float(np_state[mm.P][0, 0, 0]), damp_depth=3)
    st = {n: v.copy() for n, v in np_state.items()}
    st[mm.W] = np.zeros_like(st[mm.S])
    st["time"] = datetime(1992, 2, 20)
    return model, st


def test_moist_model_host_path_equals_oracle_numerically():
    import tasmania_b200 as tb
    from tasmania_b200.graphs import GraphedLoop
    from tasmania_b200.isentropic_moist import IsentropicMoistSUS

    nx, ny, nz, nsteps = 17, 15, 8, 4
    dt = timedelta(seconds=5)
    grid, np_state = hp.moist_case(nx, ny, nz)
    omodel, ost = _oracle_moist(nx, ny, nz, grid, np_state)
    with np.errstate(divide="ignore", invalid="ignore"):
        for _ in range(nsteps):
            ost = omodel.step(ost, dt)
        for graphed in (False, True):
            grid.topography._fact = 0.0
            with stubbed_library(OracleStub) as stub:
                FakeCapture.stub = stub
                model = IsentropicMoistSUS(grid, np_state, dt, damp_depth=3)
                if graphed:  # captured steps are replayed from the recorded calls
                    loop = GraphedLoop(model, eager_steps=1, capture_factory=FakeCapture)
                    loop.run(nsteps)
                    assert loop.period == nsteps - 1
                else:
                    model.run(nsteps)
                assert model.state["time"] == ost["time"]
                assert set(model.state) == set(ost)
                for n, v in ost.items():
                    if n != "time":
                        np.testing.assert_array_equal(tb.to_numpy(model.state[n]), v,
                                                      err_msg=f"{n} (graphed={graphed})")
    box = (slice(0, nx), slice(0, ny), slice(0, nz))
    assert float(ost[mm.QR][box].max()) > 1e-5 and float(ost[mm.ACCPREC].max()) > 0.0


def test_dry_run_host_path_equals_oracle_numerically():
    """configs[1]'s loop through the per-stencil path of the dycore mirror."""
    import tasmania_b200 as tb
    from tasmania_b200.grid import Grid, Topography, gaussian_profile, isentropic_state_from_brunt_vaisala
    from tasmania_b200.isentropic_dry import IsentropicDryRun

    nx, ny, nz, nsteps = 19, 17, 6, 5
    dt = timedelta(seconds=5)
    x, y = np.linspace(-176.0, 176.0, nx), np.linspace(-176.0, 176.0, ny)
    steady = gaussian_profile(x, y, 500.0, 50.0, 50.0)
    grid = Grid((-176.0, 176.0), nx, (-176.0, 176.0), ny, (400.0, 280.0), nz, units_to_m=1e3,
                topography=Topography(steady, timedelta(seconds=30)))
    np_state = isentropic_state_from_brunt_vaisala(grid, 22.5, 0.0, 0.015)
    pt = float(np_state[hp.P][0, 0, 0])
    ogrid = oi.Grid(nx, ny, nz, grid.dx, grid.dy, grid.dz, grid.z_on_interface_levels, grid.z)
    ohb = ob.Relaxed(nx, ny, nz, 3, 6)
    ohb.reference_state = {n: v.copy() for n, v in np_state.items()}
    otopo = hp.Topography(steady, 30.0)
    odyc = oi.IsentropicDycore(ogrid, ohb, otopo, scheme="rk3ws_si", flux="fifth_order_upwind", pt=pt,
                               eps=0.5, damp=True, damp_depth=2, damp_max=5e-4)
    ost = {n: v.copy() for n, v in np_state.items()}
    ost["time"] = datetime(2000, 1, 1)
    for step in range(nsteps):
        otopo.update((step + 1) * dt)
        out = odyc(ost, {}, dt)
        new = {n: out[n].copy() for n in (hp.S, hp.SU, hp.U, hp.SV, hp.V)}
        new["time"] = out["time"]
        for n in (hp.P, hp.EXN, hp.H, hp.MTG):
            new[n] = ost[n].copy()
        oi.refresh_diagnostics(ogrid, otopo(), new[hp.S], pt, new[hp.P], new[hp.EXN], new[hp.MTG], new[hp.H])
        ost = new
    with stubbed_library(OracleStub):
        run = IsentropicDryRun(grid, np_state, dt, damp_depth=2)
        run.dyc._fused = False  # the fused stage is one opaque ABI call: take the per-stencil path
        for _ in range(nsteps):
            run.step()
        for n in (hp.S, hp.SU, hp.SV, hp.U, hp.V, hp.MTG, hp.P, hp.EXN, hp.H):
            np.testing.assert_array_equal(tb.to_numpy(run.state[n]), ost[n], err_msg=n)
    assert float(np.abs(ost[hp.SV]).max()) > 1e-6


def test_burgers_host_path_equals_oracle_numerically():
    """configs[0]: the Burgers mirror (stepper, dycore, Dirichlet boundary with the Zhao core)."""
    import tasmania_b200 as tb
    from oracle import burgers as obu
    from tasmania_b200.boundary import Dirichlet
    from tasmania_b200.burgers import BurgersDynamicalCore, ZhaoSolutionFactory
    from tasmania_b200.grid import Grid

    nx = ny = 41
    eps, nb, nsteps = 0.01, 2, 30
    t0, dt = datetime(2000, 1, 1), timedelta(seconds=0.001)
    grid = Grid((0.0, 1.0), nx, (0.0, 1.0), ny, (0.0, 1.0), 1)
    zsf = ZhaoSolutionFactory(t0, eps)
    host = {n: np.array(zsf(t0, grid, field_name=n)) for n in ("x_velocity", "y_velocity")}
    odyc = obu.BurgersDycore(
        nx, ny, grid.dx, grid.dy, nb, scheme="rk3ws", flux="third_order",
        dirichlet=lambda time, sx, sy, name: obu.zhao_solution((time - t0).total_seconds(), grid.x[sx],
                                                                grid.y[sy], eps, name))
    ost = {n: a.copy() for n, a in host.items()}
    ost["time"] = t0
    for _ in range(nsteps):
        out = odyc(ost, {}, dt)
        ost = {"x_velocity": out["x_velocity"].copy(), "y_velocity": out["y_velocity"].copy(),
               "time": out["time"]}
    with stubbed_library(OracleStub) as stub:
        hb = Dirichlet(nx, ny, 1, nb, core=zsf, grid=grid)
        state = {n: tb.as_storage(a) for n, a in host.items()}
        state["time"] = t0
        hb.reference_state = state
        dyc = BurgersDynamicalCore(grid, hb, "rk3ws", "third_order")
        for _ in range(nsteps):
            out = dyc(state, {}, dt)
            state = {"x_velocity": out["x_velocity"].copy(), "y_velocity": out["y_velocity"].copy(),
                     "time": out["time"]}
        assert stub.count("tb200_burgers_forward_euler") == 3 * nsteps
        assert state["time"] == ost["time"]
        for n in ("x_velocity", "y_velocity"):
            np.testing.assert_array_equal(tb.to_numpy(state[n]), ost[n], err_msg=n)


def test_diffusion_dwarf_with_periodic_boundaries_equals_oracle_numerically():
    """configs[3] in miniature: fourth-order horizontal diffusion, periodic halo, phi <- phi + dt tnd."""
    import tasmania_b200 as tb
    from oracle import dwarfs
    from tasmania_b200 import stencils
    from tasmania_b200.boundary import Periodic
    from tasmania_b200.dwarfs import HorizontalDiffusion

    nx, ny, nz, nb, dt, napp = 24, 20, 6, 2, 0.05, 5
    phi0 = np.random.default_rng(20261018).standard_normal((nx, ny, nz))
    ohb = ob.Periodic(nx, ny, nz, nb)
    ophi = ohb.get_numerical_field(phi0.copy())
    shape = ophi.shape
    gamma = np.zeros(shape)
    gamma[...] = dwarfs.vertical_profile(0.5, 1.0, 3, nz)[None, None, :]
    for _ in range(napp):
        tnd = np.zeros(shape)
        dwarfs.diffusion(4, ophi, gamma, tnd, 1.0, 1.0, True, (nb, nb, 0), (nx, ny, nz))
        ophi = ophi + dt * tnd
        ohb.enforce_field(ophi)
    with stubbed_library(OracleStub) as stub:
        hb = Periodic(nx, ny, nz, nb)
        phi = tb.as_storage(np.array(hb.get_numerical_field(tb.as_storage(phi0))))
        assert phi.shape == shape
        diff = HorizontalDiffusion.factory("fourth_order", shape, 1.0, 1.0, 0.5, 1.0, 3, nb)
        tnd = tb.zeros(shape)
        for _ in range(napp):
            diff(phi, tnd, overwrite_output=True)
            stencils.fma_fields([phi], [phi], [tnd], dt, origin=(0, 0, 0), domain=shape)
            hb.enforce_field(phi)
        assert stub.count("tb200_diffusion") == napp and stub.count("tb200_periodic_enforce") >= napp
        np.testing.assert_array_equal(tb.to_numpy(phi), ophi)
        # the same loop as bench.py --workload c4 runs it
        from tasmania_b200.diffusion_dwarf import DiffusionDwarfRun

        run = DiffusionDwarfRun(nx, ny, nz, phi0, diffusion_damp_depth=3, dt=dt)
        n0 = stub.count("tb200_diffusion")
        for _ in range(napp):
            run.step()
        assert stub.count("tb200_diffusion") == n0 + napp
        np.testing.assert_array_equal(tb.to_numpy(run.phi), ophi)
        np.testing.assert_array_equal(tb.to_numpy(run.physical_field()), ophi[nb:-nb, nb:-nb])


def test_one_dimensional_dwarfs_host_path_equals_reference_fixture(stencils_1d_golden):
    """The ..._1dx / ..._1dy diffusers and smoothers through the b200 mirrors and the marshalling
    of tb200_diffusion_1d / tb200_smoothing_1d reproduce what the reference's own classes wrote
    (tests/golden/stencils_1d.npz), bit for bit."""
    import tasmania_b200 as tb
    from tasmania_b200.dwarfs import HorizontalDiffusion, HorizontalSmoothing

    fx = stencils_1d_golden
    dx, dy = fx["scalars"]
    with stubbed_library(OracleStub) as stub:
        for ax in ("x", "y"):
            for tag in ("row", "grid"):
                phi = fx[f"{ax}_{tag}_phi"]
                shape = phi.shape
                for order, name in ((2, "second_order"), (4, "fourth_order")):
                    diff = HorizontalDiffusion.factory(f"{name}_1d{ax}", shape, dx, dy, 0.5, 1.0, 3)
                    tnd = tb.zeros(shape)
                    diff(tb.as_storage(phi), tnd, overwrite_output=True)
                    np.testing.assert_array_equal(tb.to_numpy(tnd), fx[f"k8_{order}_{ax}_{tag}_tnd"])
                    acc = tb.as_storage(fx[f"{ax}_{tag}_base"])
                    diff(tb.as_storage(phi), acc, overwrite_output=False)
                    np.testing.assert_array_equal(tb.to_numpy(acc), fx[f"k8_{order}_{ax}_{tag}_acc"])
                for order, name in ((1, "first_order"), (2, "second_order"), (3, "third_order")):
                    smooth = HorizontalSmoothing.factory(f"{name}_1d{ax}", shape, 0.03, 0.24, 3)
                    out = tb.zeros(shape)
                    smooth(tb.as_storage(phi), out)
                    np.testing.assert_array_equal(tb.to_numpy(out), fx[f"k9_{order}_{ax}_{tag}_out"])
        assert stub.count("tb200_diffusion_1d") == 16 and stub.count("tb200_smoothing_1d") == 12
        assert stub.count("tb200_diffusion") == 0 and stub.count("tb200_smoothing") == 0


def test_thomas_stencil_host_path_equals_reference_fixture(stencils_1d_golden):
    """compile_stencil("thomas") -> tb200_thomas marshalling (sliced, non-contiguous views
    included) reproduces the reference's thomas_numpy output."""
    import tasmania_b200 as tb
    from tasmania_b200.framework import BackendOptions

    fx = stencils_1d_golden
    box = [int(v) for v in fx["thomas_box"]]
    with stubbed_library(OracleStub) as stub:
        thomas = tb.compile_stencil("thomas", backend_options=BackendOptions())
        a, b, c, d = (tb.as_storage(fx["thomas_" + n]) for n in "abcd")
        x = tb.zeros(fx["thomas_a"].shape)
        thomas(a=a, b=b, c=c, d=d, out=x, origin=tuple(box[:3]), domain=tuple(box[3:]))
        np.testing.assert_array_equal(tb.to_numpy(x), fx["thomas_x"])
        # the same systems through views of the storages, solved in place of d
        sl = (slice(1, 6), slice(0, 5), slice(1, 8))
        dd = tb.as_storage(fx["thomas_d"])
        thomas(a=a[sl], b=b[sl], c=c[sl], d=dd[sl], out=dd[sl], origin=(0, 0, 0), domain=(5, 5, 7))
        np.testing.assert_array_equal(tb.to_numpy(dd)[sl], fx["thomas_x"][sl])
        assert stub.count("tb200_thomas") == 2


def test_hyperdiffusion_stencil_host_path_equals_reference_fixture(stencils_1d_golden):
    import tasmania_b200 as tb
    from tasmania_b200.framework import BackendOptions

    fx = stencils_1d_golden
    box = [int(v) for v in fx["hyper_box"]]
    with stubbed_library(OracleStub) as stub:
        hyper = tb.compile_stencil("hyperdiffusion", backend_options=BackendOptions())
        out = tb.zeros(fx["hyper_phi"].shape)
        hyper(in_phi=tb.as_storage(fx["hyper_phi"]), out_phi=out, alpha=float(fx["hyper_alpha"]),
              origin=tuple(box[:3]), domain=tuple(box[3:]))
        np.testing.assert_array_equal(tb.to_numpy(out), fx["hyper_out"])
        assert stub.count("tb200_hyperdiffusion") == 1


def test_one_dimensional_boundary_mirrors_equal_reference_fixture(stencils_1d_golden):
    import tasmania_b200 as tb

    with stubbed_library(OracleStub) as stub:
        hp.check_one_dimensional_boundaries(stencils_1d_golden, tb)
        assert stub.count("tb200_relax") == 8


def test_fused_stage_host_path_equals_oracle_numerically():
    """The headline path's host side: IsentropicDryRun with the fused stage (one ABI call per RK
    stage carrying 25 fields and the stage configuration), emulated stage by stage with the oracle."""
    import tasmania_b200 as tb
    from tasmania_b200.graphs import GraphedLoop
    from tasmania_b200.grid import Grid, Topography, gaussian_profile, isentropic_state_from_brunt_vaisala
    from tasmania_b200.isentropic_dry import IsentropicDryRun

    nx, ny, nz, nsteps = 19, 17, 6, 5
    dt = timedelta(seconds=5)
    x, y = np.linspace(-176.0, 176.0, nx), np.linspace(-176.0, 176.0, ny)
    steady = gaussian_profile(x, y, 500.0, 50.0, 50.0)

    def case():
        grid = Grid((-176.0, 176.0), nx, (-176.0, 176.0), ny, (400.0, 280.0), nz, units_to_m=1e3,
                    topography=Topography(steady, timedelta(seconds=30)))
        return grid, isentropic_state_from_brunt_vaisala(grid, 22.5, 0.0, 0.015)

    results = {}
    for mode in ("stencils", "fused", "fused+graphs"):
        grid, np_state = case()
        with stubbed_library(OracleStub) as stub:
            FakeCapture.stub = stub
            run = IsentropicDryRun(grid, np_state, dt, damp_depth=2)
            assert run.dyc._fused
            if mode == "stencils":
                run.dyc._fused = False
            stepper = GraphedLoop(run, eager_steps=1, capture_factory=FakeCapture) if "graphs" in mode else run
            for _ in range(nsteps):
                stepper.step()
            if mode == "fused":
                assert stub.count("tb200_isentropic_stage_dry") == 3 * nsteps
            if mode == "fused+graphs":  # two buffer configurations captured, the rest replayed
                assert stepper.period == 2 and stepper.replayed_launches >= 4 * (nsteps - 1)
            results[mode] = {n: tb.to_numpy(v) for n, v in run.state.items() if n != "time"}
    for mode in ("fused", "fused+graphs"):
        for n, v in results["stencils"].items():
            np.testing.assert_array_equal(results[mode][n], v, err_msg=f"{mode}: {n}")
    assert float(np.abs(results["fused"][hp.SV]).max()) > 1e-6


def test_decomposed_run_host_path_equals_single_domain_numerically():
    """SURVEY.md section 8e on the host: a 2 x 2 (and 3 x 1) decomposition stepped sub-domain by sub-domain
    with halo exchanges equals the single-domain run bit for bit (kernels emulated by the oracle)."""
    from tasmania_b200.distributed import InProcessDecomposedRun

    nxg, nyg, nz, nsteps = 37, 31, 5, 3
    kw = dict(damp_depth=2, topo_seconds=15.0, device="cpu")
    with stubbed_library(OracleStub):
        single = InProcessDecomposedRun(nxg, nyg, nz, 1, 1, **kw)
        for _ in range(nsteps):
            single.step()
        want = {n: single.gather(n) for n in single.subs[0].names}
        for px, py in ((2, 2), (3, 1)):
            run = InProcessDecomposedRun(nxg, nyg, nz, px, py, **kw)
            for _ in range(nsteps):
                run.step()
            for n, v in want.items():
                np.testing.assert_array_equal(run.gather(n), v, err_msg=f"{px}x{py}: {n}")
    assert float(np.abs(want["y_momentum_isentropic"]).max()) > 1e-6


def test_fused_stage_with_slow_tendencies_equals_stencil_path_numerically():
    """VERDICT round 1, missing 4: slow tendencies of s, su, sv passed to the dycore ride the fused
    stage (tb200_isentropic_stage.s_tnd / su_tnd / sv_tnd) instead of dropping to the per-stencil
    path; a subset of the three is completed by a field of zeros.  Same bits as the stencil path
    (K1 / K2 with tendencies, prognostics/utils.py:L43-L204), here through the oracle-backed stub."""
    import tasmania_b200 as tb
    from tasmania_b200.boundary import Relaxed
    from tasmania_b200.grid import Grid, Topography, gaussian_profile, isentropic_state_from_brunt_vaisala
    from tasmania_b200.isentropic import IsentropicDynamicalCore

    nx, ny, nz, nb = 19, 17, 6, 3
    dt = timedelta(seconds=5)
    x, y = np.linspace(-176.0, 176.0, nx), np.linspace(-176.0, 176.0, ny)
    rng = np.random.default_rng(3)
    shape = (nx + 1, ny + 1, nz + 1)
    tnd_np = {hp.S: 1e-3 * rng.standard_normal(shape), hp.SU: 1e-2 * rng.standard_normal(shape),
              hp.SV: 1e-2 * rng.standard_normal(shape)}
    results = {}
    for mode, keys in (("stencils", (hp.S, hp.SU, hp.SV)), ("fused", (hp.S, hp.SU, hp.SV)),
                       ("stencils-subset", (hp.SU,)), ("fused-subset", (hp.SU,))):
        grid = Grid((-176.0, 176.0), nx, (-176.0, 176.0), ny, (400.0, 280.0), nz, units_to_m=1e3,
                    topography=Topography(gaussian_profile(x, y, 500.0, 50.0, 50.0), timedelta(seconds=30)))
        np_state = isentropic_state_from_brunt_vaisala(grid, 22.5, 0.0, 0.015)
        with stubbed_library(OracleStub) as stub:
            state = {n: tb.as_storage(v) for n, v in np_state.items()}
            state["time"] = datetime(2000, 1, 1)
            hb = Relaxed(nx, ny, nz, nb, nr=6)
            hb.reference_state = state
            dyc = IsentropicDynamicalCore(
                grid, hb, time_integration_scheme="rk3ws_si", horizontal_flux_scheme="fifth_order_upwind",
                time_integration_properties={"pt": float(np_state[hp.P][0, 0, 0]), "eps": 0.5}, damp=True,
                damp_depth=2, damp_max=5e-4, fused=mode.startswith("fused"))
            tnd = {n: tb.as_storage(tnd_np[n]) for n in keys}
            dyc.update_topography(dt)
            out = dyc(state, tnd, dt)
            fused_calls = stub.count("tb200_isentropic_stage_dry")
            assert fused_calls == (3 if mode.startswith("fused") else 0), (mode, fused_calls)
            results[mode] = {n: tb.to_numpy(out[n]) for n in (hp.S, hp.SU, hp.SV, hp.U, hp.V)}
    for a, b in (("stencils", "fused"), ("stencils-subset", "fused-subset")):
        for n, v in results[a].items():
            np.testing.assert_array_equal(results[b][n][:nx, :ny, :nz], v[:nx, :ny, :nz], err_msg=f"{b}: {n}")
    assert not np.array_equal(results["fused"][hp.SU], results["fused-subset"][hp.SU])


def periodic_case(nx, ny, nz, nb, seed=5):
    """A developed state on the NUMERICAL grid (nx, ny) of a periodic domain: hydrostatic background,
    perturbed s, su, sv wrapped into the nb ghost layers (periodic.py:L98-L122), u, v diagnosed."""
    from oracle import dwarfs
    from tasmania_b200.grid import Grid, Topography, gaussian_profile, isentropic_state_from_brunt_vaisala

    x, y = np.linspace(-176.0, 176.0, nx), np.linspace(-176.0, 176.0, ny)
    steady = gaussian_profile(x, y, 500.0, 50.0, 50.0)
    grid = Grid((-176.0, 176.0), nx, (-176.0, 176.0), ny, (400.0, 280.0), nz, units_to_m=1e3,
                topography=Topography(steady, timedelta(seconds=30)))
    st = isentropic_state_from_brunt_vaisala(grid, 22.5, 0.0, 0.015)
    rng = np.random.default_rng(seed)
    ohb = ob.Periodic(nx - 2 * nb, ny - 2 * nb, nz, nb)
    for n, amp in ((hp.S, 0.02), (hp.SU, 0.05), (hp.SV, 0.05)):
        st[n] = st[n] * (1.0 + amp * rng.standard_normal(st[n].shape))
        if n == hp.SV:
            st[n] = st[n] + 5.0 * st[hp.S] * rng.standard_normal(st[n].shape)
        ohb.enforce_field(st[n], n)
    dwarfs.get_velocity_components(nx, ny, nz, st[hp.S], st[hp.SU], st[hp.SV], st[hp.U], st[hp.V])
    ohb.set_outermost_layers_x(st[hp.U], hp.U)
    ohb.set_outermost_layers_y(st[hp.V], hp.V)
    return grid, steady, st


def oracle_periodic_run(grid, steady, np_state, nb, nsteps, dt, tendencies=None, damp=True):
    nx, ny, nz = grid.nx, grid.ny, grid.nz
    pt = float(np_state[hp.P][0, 0, 0])
    ogrid = oi.Grid(nx, ny, nz, grid.dx, grid.dy, grid.dz, grid.z_on_interface_levels, grid.z)
    ohb = ob.Periodic(nx - 2 * nb, ny - 2 * nb, nz, nb)
    ohb.reference_state = {n: v.copy() for n, v in np_state.items()}
    otopo = hp.Topography(steady, 30.0)
    odyc = oi.IsentropicDycore(ogrid, ohb, otopo, scheme="rk3ws_si", flux="fifth_order_upwind", pt=pt,
                               eps=0.5, damp=damp, damp_depth=2, damp_max=5e-4)
    ost = {n: v.copy() for n, v in np_state.items()}
    ost["time"] = datetime(2000, 1, 1)
    for step in range(nsteps):
        otopo.update((step + 1) * dt)
        out = odyc(ost, tendencies or {}, dt)
        new = {n: out[n].copy() for n in (hp.S, hp.SU, hp.U, hp.SV, hp.V)}
        new["time"] = out["time"]
        for n in (hp.P, hp.EXN, hp.H, hp.MTG):
            new[n] = ost[n].copy()
        oi.refresh_diagnostics(ogrid, otopo(), new[hp.S], pt, new[hp.P], new[hp.EXN], new[hp.MTG], new[hp.H])
        ost = new
    return ost


def b200_periodic_run(tb, grid, np_state, nb, nsteps, dt, fused, tendencies=None, damp=True, lazy=True):
    """The dry loop of tasmania_b200.isentropic_dry with a Periodic boundary (that class builds a Relaxed one)."""
    from tasmania_b200.boundary import Periodic
    from tasmania_b200.isentropic import IsentropicDiagnostics, IsentropicDynamicalCore

    nx, ny, nz = grid.nx, grid.ny, grid.nz
    pt = float(np_state[hp.P][0, 0, 0])
    hb = Periodic(nx - 2 * nb, ny - 2 * nb, nz, nb)
    state = {n: tb.as_storage(v) for n, v in np_state.items()}
    state["time"] = datetime(2000, 1, 1)
    hb.reference_state = {n: tb.as_storage(v) for n, v in np_state.items()}
    dyc = IsentropicDynamicalCore(
        grid, hb, time_integration_scheme="rk3ws_si", horizontal_flux_scheme="fifth_order_upwind",
        time_integration_properties={"pt": pt, "eps": 0.5}, damp=damp, damp_depth=2, damp_max=5e-4, fused=fused)
    assert dyc._fused == fused and (not fused or dyc._periodic)
    dyc.lazy_velocities = bool(fused and lazy)
    diag = IsentropicDiagnostics(grid)
    tnd = {n: tb.as_storage(v) for n, v in (tendencies or {}).items()}
    for step in range(nsteps):
        dyc.update_topography((step + 1) * dt)
        dyc._prognostic._diagnostics._set_topography()
        diag._set_topography()
        out = dyc(state, tnd, dt)
        new = {n: out[n] for n in (hp.S, hp.SU, hp.U, hp.SV, hp.V)}
        new["time"] = out["time"]
        for n in (hp.P, hp.EXN, hp.H, hp.MTG):
            new[n] = tb.zeros(dyc.storage_shape)
        diag.get_diagnostic_variables(new[hp.S], pt, new[hp.P], new[hp.EXN], new[hp.MTG], new[hp.H])
        state = new
    return {n: tb.to_numpy(v) for n, v in state.items() if n != "time"}


def test_periodic_dry_dycore_fused_and_stencil_paths_equal_oracle_numerically():
    """VERDICT round 1, missing 4: the dry core with a Periodic boundary takes the fused stage too
    (tb200_isentropic_stage.periodic: gamma = 0, the wrap of s between the s-step and the scans inside
    the call, enforce_raw and the damping after it in the reference's order, dycore.py:L684-L700).
    Per-stencil path, fused path, fused path with slow tendencies: the oracle dycore's bits."""
    import tasmania_b200 as tb

    nx, ny, nz, nb, nsteps = 21, 19, 6, 3, 3
    dt = timedelta(seconds=5)
    grid, steady, np_state = periodic_case(nx, ny, nz, nb)
    rng = np.random.default_rng(11)
    tnd = {hp.SU: 1e-2 * rng.standard_normal(np_state[hp.SU].shape)}
    for tendencies in (None, tnd):
        want = oracle_periodic_run(grid, steady, np_state, nb, nsteps, dt, tendencies)
        for fused, lazy in ((False, False), (True, True), (True, False)):
            grid, steady, _ = periodic_case(nx, ny, nz, nb)
            with stubbed_library(OracleStub) as stub:
                got = b200_periodic_run(tb, grid, np_state, nb, nsteps, dt, fused, tendencies, lazy=lazy)
                assert stub.count("tb200_isentropic_stage_dry") == (3 * nsteps if fused else 0)
                assert stub.count("tb200_periodic_enforce") >= 4 * 3 * nsteps - (3 * nsteps if fused else 0)
            for n in (hp.S, hp.SU, hp.SV, hp.U, hp.V, hp.MTG):
                np.testing.assert_array_equal(got[n][:nx + 1, :ny + 1, :nz], want[n][:nx + 1, :ny + 1, :nz],
                                              err_msg=f"fused={fused} lazy={lazy} tendencies={tendencies is not None}: {n}")
        assert float(np.abs(want[hp.SV] - np_state[hp.SV]).max()) > 1e-6


@pytest.mark.parametrize("case", ("isen_dry_rk3_5th_periodic", "isen_dry_rk3_3rd_periodic",
                                  "isen_dry_fe_cen_periodic", "isen_moist_rk3_5th_periodic"))
@pytest.mark.parametrize("fused,lazy", ((False, False), (True, True), (True, False)))
def test_periodic_dycore_host_path_equals_reference_fixture(case, fused, lazy):
    """The harness of tests/test_gpu_isentropic.py over the oracle-backed stub: the dycore mirror with
    the Periodic boundary -- per-stencil path and fused stage (dry and moist) -- reproduces, bit for
    bit, the fixtures written by the reference's own ``stage_array_call_dry / _moist`` on a periodic
    domain."""
    import tasmania_b200 as tb
    from tests import test_gpu_isentropic as tg

    fx = hp.load(case)
    with stubbed_library(OracleStub) as stub:
        grid, hb, dyc, diag, state, pt, dt, nsteps = tg.build_from_fixture(fx, fused)
        assert dyc._fused == fused and (not fused or dyc._periodic)
        dyc.lazy_velocities = lazy
        final, stage0 = tg.run(grid, dyc, diag, state, pt, dt, nsteps)
        entry = "tb200_isentropic_stage_moist" if dyc._moist else "tb200_isentropic_stage_dry"
        assert stub.count(entry) == (dyc.stages * nsteps if fused else 0)
        nx, ny, nz = grid.nx, grid.ny, grid.nz
        qn = tg.QNAMES if dyc._moist else ()
        for n in (hp.S, hp.SU, hp.SV) + qn + (() if lazy else (hp.U, hp.V)):
            np.testing.assert_array_equal(stage0[n][: nx + 1, : ny + 1, :nz],
                                          fx["stage0_" + n][: nx + 1, : ny + 1, :nz], err_msg="stage0 " + n)
        for n in (hp.S, hp.SU, hp.SV, hp.U, hp.V, hp.MTG, hp.P, hp.EXN, hp.H) + qn:
            np.testing.assert_array_equal(tb.to_numpy(final[n])[: nx + 1, : ny + 1, : nz + 1],
                                          fx["final_" + n][: nx + 1, : ny + 1, : nz + 1], err_msg="final " + n)


def test_dry_run_with_periodic_boundaries_equals_oracle_numerically():
    """IsentropicDryRun(boundary="periodic"): the benchmark loop (dycore with the fused periodic
    stage -> diagnostics refresh, ping-pong buffers), eagerly and replayed from captured graphs,
    against the oracle's periodic dycore loop -- bit for bit through the oracle-backed stub."""
    import tasmania_b200 as tb
    from tasmania_b200.graphs import GraphedLoop
    from tasmania_b200.isentropic_dry import IsentropicDryRun

    nx, ny, nz, nb, nsteps = 21, 19, 6, 3, 4
    dt = timedelta(seconds=5)
    grid, steady, np_state = periodic_case(nx, ny, nz, nb)
    want = oracle_periodic_run(grid, steady, np_state, nb, nsteps, dt)
    for mode in ("eager", "graphs"):
        grid, steady, _ = periodic_case(nx, ny, nz, nb)
        with stubbed_library(OracleStub) as stub:
            FakeCapture.stub = stub
            run = IsentropicDryRun(grid, np_state, dt, damp_depth=2, boundary="periodic")
            assert run.dyc._fused and run.dyc._periodic
            stepper = GraphedLoop(run, eager_steps=1, capture_factory=FakeCapture) if mode == "graphs" else run
            for _ in range(nsteps):
                stepper.step()
            assert stub.count("tb200_isentropic_stage_dry") >= 3 * (nsteps if mode == "eager" else 2)
            for n in (hp.S, hp.SU, hp.SV, hp.U, hp.V, hp.MTG, hp.P, hp.EXN, hp.H):
                np.testing.assert_array_equal(tb.to_numpy(run.state[n]), want[n], err_msg=f"{mode}: {n}")
