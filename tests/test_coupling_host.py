# -*- coding: utf-8 -*-
"""Host logic of tasmania_b200.coupling and of the moist-model assembly, against the recording ABI
stub of tests/abi_stub.py (no GPU, no kernels: only ``tb200_fma_fields`` and copy are carried
out on host buffers).  Numbers are compared with the oracle's straight-line ``tendency_step``."""
from datetime import datetime, timedelta

import numpy as np
import pytest

from oracle import moist_model as mm
from tests import helpers as hp
from tests.abi_stub import stubbed_library


class Decay:
    """dy/dt = -rate * y as a tendency component working on host storages (test double)."""

    kind, diagnostic_names = "tendency", ("seen",)

    def __init__(self, rate, names=("y",)):
        self.rate, self.tendency_names = rate, tuple(names)

    def diagnostic_shape(self, name):
        return (4, 3, 2)

    def array_call(self, state, out_tendencies, out_diagnostics, overwrite_tendencies):
        for n in self.tendency_names:
            t = -self.rate * state[n].t
            if overwrite_tendencies[n]:
                out_tendencies[n].t.copy_(t)
            else:
                out_tendencies[n].t.add_(t)
        out_diagnostics["seen"].t.copy_(state[self.tendency_names[0]].t)


@pytest.mark.parametrize("scheme", ["forward_euler", "rk2", "rk3ws"])
def test_tendency_stepper_matches_oracle_scheme(scheme):
    import tasmania_b200 as tb
    from tasmania_b200.coupling import TendencyStepper

    rng = np.random.default_rng(3)
    y0, z0 = rng.standard_normal((4, 3, 2)), rng.standard_normal((4, 3, 2))
    dt = timedelta(seconds=0.3)
    with stubbed_library() as stub:
        state = {"y": tb.as_storage(y0), "z": tb.as_storage(z0), "other": tb.as_storage(y0),
                 "time": datetime(2000, 1, 1)}
        # two components on y: the second one accumulates (overwrite False)
        stepper = TendencyStepper.factory(scheme, Decay(1.0, ("y", "z")), Decay(0.5, ("y",)))
        diags, out = stepper(state, dt)
        got = {n: tb.to_numpy(out[n]) for n in ("y", "z")}
        seen = tb.to_numpy(diags["seen"])
        nstages = len(TendencyStepper.SCHEMES[scheme])
        assert stub.count("tb200_fma_fields") == nstages  # ONE launch per stage for both fields
        assert out["time"] == datetime(2000, 1, 1) + dt and set(out) == {"y", "z", "time"}

    def fn(st):
        return {"y": -1.0 * st["y"] + -0.5 * st["y"], "z": -1.0 * st["z"]}, {"seen": st["y"].copy()}

    _, want = mm.tendency_step(scheme, {"y": y0, "z": z0, "other": y0}, fn, dt.total_seconds())
    np.testing.assert_array_equal(got["y"], want["y"])
    np.testing.assert_array_equal(got["z"], want["z"])
    np.testing.assert_array_equal(seen, y0)  # diagnostics of the first stage


def test_promoters_and_overwrite_flags():
    """d2t copies the state's diagnostic into the tendency buffer, the next component
    accumulates on top, t2d promotes the sum (driver_namelist_sus.py:L320-L366)."""
    import tasmania_b200 as tb
    from tasmania_b200.coupling import (AirPotentialTemperatureToDiagnostic,
                                        AirPotentialTemperatureToTendency, ConcurrentCoupling)
    from tasmania_b200.grid import Grid

    class Heating:
        kind, diagnostic_names = "tendency", ()
        tendency_names = ("air_potential_temperature",)
        seen_overwrite = None

        def array_call(self, state, out_tendencies, out_diagnostics, overwrite_tendencies):
            self.seen_overwrite = dict(overwrite_tendencies)
            assert not overwrite_tendencies["air_potential_temperature"]
            out_tendencies["air_potential_temperature"].t.add_(2.0)

    grid = Grid((0.0, 1.0), 3, (0.0, 1.0), 2, (300.0, 280.0), 1)
    w0 = np.arange(24.0).reshape(4, 3, 2)
    with stubbed_library():
        state = {"tendency_of_air_potential_temperature": tb.as_storage(w0),
                 "air_isentropic_density": tb.as_storage(w0)}
        heat = Heating()
        cc = ConcurrentCoupling(AirPotentialTemperatureToTendency(grid), heat,
                                AirPotentialTemperatureToDiagnostic(grid))
        assert cc.overwrite_tendencies == [{"air_potential_temperature": True},
                                           {"air_potential_temperature": False}, {}]
        tnd, diag = cc(state, timedelta(seconds=1))
        got = tb.to_numpy(diag["tendency_of_air_potential_temperature"])
    want = np.zeros_like(w0)
    want[:3, :2, :1] = w0[:3, :2, :1] + 2.0  # the promoters copy the grid box only
    np.testing.assert_array_equal(got, want)


def test_sequential_update_splitting_swaps_buffers():
    import tasmania_b200 as tb
    from tasmania_b200.coupling import SequentialUpdateSplitting, TimeIntegrationOptions

    class Doubler:
        kind, tendency_names, diagnostic_names = "diagnostic", (), ("y",)

        def diagnostic_shape(self, name):
            return (4, 3, 2)

        def zeros(self, *, shape):
            return tb.zeros(shape)

        def array_call(self, state, out):
            out["y"].t.copy_(2.0 * state["y"].t)

    y0 = np.ones((4, 3, 2))
    dt = timedelta(seconds=0.5)
    with stubbed_library():
        state = {"y": tb.as_storage(y0), "time": datetime(2000, 1, 1)}
        first = state["y"]
        sus = SequentialUpdateSplitting(TimeIntegrationOptions(Doubler()),
                                        TimeIntegrationOptions(Decay(1.0), scheme="forward_euler"))
        sus(state, dt)
        np.testing.assert_array_equal(tb.to_numpy(state["y"]), 2.0 * y0 * (1 - 0.5))
        assert state["time"] == datetime(2000, 1, 1) + dt
        # the array that held y went back to the component as its next output buffer
        assert sus._out_diagnostics[0]["y"] is first
        sus(state, dt)
        np.testing.assert_array_equal(tb.to_numpy(state["y"]), 4.0 * y0 * (1 - 0.5) ** 2)


def test_moist_model_host_path_end_to_end():
    """The whole configs[2] assembly (dycore + ten physics entries) runs through the ABI stub:
    every kernel call type-checks against the declared C signature, the sequence of launches of a
    step is the one the driver's component list implies, and buffers are recycled (no allocation
    after the second step)."""
    from tasmania_b200.isentropic_moist import IsentropicMoistSUS

    nx, ny, nz = 17, 15, 8
    grid, np_state = hp.moist_case(nx, ny, nz)
    with stubbed_library() as stub:
        model = IsentropicMoistSUS(grid, np_state, timedelta(seconds=5), damp_depth=3)
        model.step()
        calls = list(stub.calls)
        per_step = {n: calls.count(n) for n in set(calls)}
        # physics stages: Coriolis rk2 (2), Smagorinsky rk2 (2), Kessler rk2 (2), saturation rk2
        # (2), vertical advection rk3ws (3), sedimentation rk3ws (3)
        # (Coriolis and Smagorinsky apply their stage update themselves: no tendencies + fma round trip)
        assert per_step["tb200_coriolis_step"] == 2 and per_step["tb200_smagorinsky_step"] == 2
        assert "tb200_coriolis" not in per_step and "tb200_smagorinsky" not in per_step
        assert per_step["tb200_kessler"] == 2 and per_step["tb200_saturation_prognostic"] == 2
        # (the vertical advection applies its stage update itself: no tendencies + fma round trip)
        assert per_step["tb200_vertical_advection_step"] == 3 and per_step["tb200_sedimentation"] == 3
        assert "tb200_vertical_advection" not in per_step
        assert per_step["tb200_fall_velocity"] == 4          # 3 sedimentation stages + precipitation
        assert per_step["tb200_accumulated_precipitation"] == 1
        assert per_step["tb200_smoothing"] == 6              # s, su, sv, qv, qc, qr
        assert per_step["tb200_fma_fields"] == 7             # one per remaining stepper stage
        assert per_step["tb200_diagnostic_variables"] == 1
        assert per_step["tb200_density_and_temperature"] == 1
        # moist dycore: one fused call per RK stage (s-step, tracers, scans, momentum, relaxation,
        # damping) and one velocity pass for the step's final state -- none of the reference's
        # per-stencil launches
        assert per_step["tb200_isentropic_stage_moist"] == 3 and per_step["tb200_velocity_components"] == 1
        assert not {"tb200_step_forward_euler", "tb200_step_forward_euler_momentum", "tb200_relax_frame",
                    "tb200_density", "tb200_mass_fraction", "tb200_damping"} & set(per_step)
        # order: dynamics first, then the physics in the driver's order
        first = {n: calls.index(n) for n in per_step}
        order = ["tb200_isentropic_stage_moist", "tb200_diagnostic_variables", "tb200_coriolis_step",
                 "tb200_smoothing", "tb200_smagorinsky_step", "tb200_kessler",
                 "tb200_saturation_prognostic", "tb200_vertical_advection_step", "tb200_sedimentation",
                 "tb200_accumulated_precipitation"]
        assert [first[n] for n in order] == sorted(first[n] for n in order)
        names1 = set(model.state)
        model.step()
        ids2 = {id(v.t) for n, v in model.state.items() if n != "time"}
        from tasmania_b200 import storage

        n_alloc = [0]
        real_allocate = storage._allocate

        def counting_allocate(*a, **k):
            n_alloc[0] += 1
            return real_allocate(*a, **k)

        storage._allocate = counting_allocate
        try:
            model.step()
            model.step()
        finally:
            storage._allocate = real_allocate
        assert n_alloc[0] == 0, "steady-state steps must not allocate"
        assert set(model.state) == names1
        assert len(ids2) == len(names1) - 1                  # no array sits under two names
        assert model.state["time"] == datetime(1992, 2, 20) + 4 * timedelta(seconds=5)


def test_fused_stage_update_alternates_stage_buffers(monkeypatch):
    """A lone component with ``array_call_stepped`` replaces tendencies + fma; every stage reads
    the previous stage's output and the last one writes ``out_state``; TB200_FUSED_STEP=0 turns
    the fusion off.  Checked on numbers against the oracle scheme."""
    import tasmania_b200 as tb
    from tasmania_b200.coupling import TendencyStepper

    class FusableDecay(Decay):
        diagnostic_names = ()
        stepped_calls = 0

        def array_call(self, state, out_tendencies, out_diagnostics, overwrite_tendencies):
            for n in self.tendency_names:
                out_tendencies[n].t.copy_(-self.rate * state[n].t)

        def array_call_stepped(self, state, base, factor, out_state):
            self.stepped_calls += 1
            for n in self.tendency_names:
                assert out_state[n] is not state[n] and out_state[n] is not base[n]
                out_state[n].t.copy_(base[n].t + factor * (-self.rate * state[n].t))

    y0 = np.random.default_rng(9).standard_normal((4, 3, 2))
    dt = timedelta(seconds=0.3)
    for scheme in ("forward_euler", "rk2", "rk3ws"):
        _, want = mm.tendency_step(scheme, {"y": y0}, lambda st: ({"y": -0.7 * st["y"]}, {}),
                                   dt.total_seconds())
        for fused in (True, False):
            monkeypatch.setenv("TB200_FUSED_STEP", "1" if fused else "0")
            with stubbed_library() as stub:
                comp = FusableDecay(0.7)
                stepper = TendencyStepper.factory(scheme, comp)
                state = {"y": tb.as_storage(y0)}
                _, out = stepper(state, dt)
                nst = len(TendencyStepper.SCHEMES[scheme])
                assert comp.stepped_calls == (nst if fused else 0)
                assert stub.count("tb200_fma_fields") == (0 if fused else nst)
                np.testing.assert_array_equal(tb.to_numpy(out["y"]), want["y"])
                np.testing.assert_array_equal(tb.to_numpy(state["y"]), y0)


def test_as_parallel_policy_components_do_not_see_each_other():
    """``as_parallel`` (concurrent_coupling.py:L376-L423): every component gets the input state,
    diagnostics of one are not visible to the next, tendency -> diagnostic promoters are skipped."""
    import tasmania_b200 as tb
    from tasmania_b200.coupling import ConcurrentCoupling, FromTendencyToDiagnostic
    from tasmania_b200.grid import Grid

    class Diag:
        kind, tendency_names, diagnostic_names = "diagnostic", (), ("d",)

        def diagnostic_shape(self, name):
            return (4, 3, 2)

        def array_call(self, state, out):
            out["d"].t.copy_(state["y"].t + 1.0)

    class UsesD:
        kind, tendency_names, diagnostic_names = "tendency", ("y",), ()

        def array_call(self, state, out_tendencies, out_diagnostics, overwrite_tendencies):
            out_tendencies["y"].t.copy_(state["d"].t)

    y0 = np.ones((4, 3, 2))
    grid = Grid((0.0, 1.0), 3, (0.0, 1.0), 2, (300.0, 280.0), 1)
    for policy, want in (("serial", 2.0), ("as_parallel", 7.0)):
        with stubbed_library():
            state = {"y": tb.as_storage(y0), "d": tb.as_storage(7.0 * y0)}
            cc = ConcurrentCoupling(Diag(), UsesD(), FromTendencyToDiagnostic(grid, "y"),
                                    execution_policy=policy)
            tnd, diag = cc(state, timedelta(seconds=1))
            np.testing.assert_array_equal(tb.to_numpy(tnd["y"]), want * y0)
            promoted = tb.to_numpy(diag["tendency_of_y"])
            assert (promoted[:3, :2, :1] == (want if policy == "serial" else 0.0)).all()
    with pytest.raises(ValueError):
        ConcurrentCoupling(Diag(), execution_policy="concurrent")


def test_fused_stage_is_for_the_two_dimensional_relaxed_boundary_only():
    """Relaxed1DX / Relaxed1DY also report type "relaxed", but the fused kernels
    assume the 2-D class (gamma == 1 on the nb outer rings, no repetition of the middle line): the
    dycore must not select the fused stage for them, and must refuse it when forced."""
    from tasmania_b200.boundary import HorizontalBoundary, Relaxed, Relaxed1DX
    from tasmania_b200.grid import Grid
    from tasmania_b200.isentropic import IsentropicDynamicalCore

    with stubbed_library():
        nx, nz, nb = 21, 6, 3
        hb1 = HorizontalBoundary.factory("relaxed", nx, 1, nz, nb, nr=6)
        assert isinstance(hb1, Relaxed1DX) and hb1.type == "relaxed"
        grid1 = Grid((0.0, 100.0), hb1.ni, (0.0, 100.0), hb1.nj, (340.0, 280.0), nz)
        kw = dict(time_integration_scheme="rk3ws_si", horizontal_flux_scheme="fifth_order_upwind",
                  time_integration_properties={"pt": 1000.0, "eps": 0.5}, damp=False)
        dyc = IsentropicDynamicalCore(grid1, hb1, **kw)
        assert not dyc._fused and not dyc.lazy_velocities
        with pytest.raises(ValueError):
            IsentropicDynamicalCore(grid1, hb1, fused=True, **kw)
        hb2 = HorizontalBoundary.factory("relaxed", nx, 19, nz, nb, nr=6)
        assert type(hb2) is Relaxed
        grid2 = Grid((0.0, 100.0), nx, (0.0, 100.0), 19, (340.0, 280.0), nz)
        assert IsentropicDynamicalCore(grid2, hb2, **kw)._fused
