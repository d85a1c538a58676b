# -*- coding: utf-8 -*-
"""tasmania_b200.graphs.GraphedLoop on the host: graph capture is played by tests/abi_stub.py's
FakeCapture (records ABI calls without executing them, re-issues them on replay), which is all
that is needed to check the part that can go wrong -- the keying of graphs by buffer configuration
and the restoration of the model's dict bookkeeping after a replay."""
from datetime import timedelta

import numpy as np

from tests import helpers as hp
from tests.abi_stub import FakeCapture, canonical, stubbed_library


class Rotating:
    """y <- y + f * inc through three rotating buffers plus a ping-pong pair (period 6)."""

    def __init__(self, tb, stencils, y0):
        self.stencils = stencils
        self.state = {"y": tb.as_storage(y0), "z": tb.as_storage(2.0 * y0)}
        self.pool = {"a": tb.zeros(y0.shape), "b": tb.zeros(y0.shape)}
        self.pong = {"z": tb.zeros(y0.shape)}
        self.inc = tb.as_storage(np.ones_like(y0))
        self.nstep = 0

    def prepare_step(self):
        self.nstep += 1

    def compute_step(self):
        st, pool, pong = self.state, self.pool, self.pong
        shape = st["y"].shape
        self.stencils.fma_fields([pool["a"], pong["z"]], [st["y"], st["z"]], [self.inc, st["y"]], 0.5,
                                 origin=(0, 0, 0), domain=shape)
        st["y"], pool["a"], pool["b"] = pool["a"], pool["b"], st["y"]  # 3-cycle
        st["z"], pong["z"] = pong["z"], st["z"]                         # 2-cycle

    def finish_step(self):
        pass

    def buffer_dicts(self):
        return [self.state, self.pool, self.pong]

    def set_buffer_dicts(self, dicts):
        self.state, self.pool, self.pong = dicts


def test_graphed_loop_equals_eager_on_rotating_buffers():
    import tasmania_b200 as tb
    from tasmania_b200 import stencils
    from tasmania_b200.graphs import GraphedLoop

    y0 = np.random.default_rng(5).standard_normal((5, 4, 3))
    with stubbed_library() as stub:
        FakeCapture.stub = stub
        eager, graphed = Rotating(tb, stencils, y0), Rotating(tb, stencils, y0)
        loop = GraphedLoop(graphed, eager_steps=1, capture_factory=FakeCapture)
        for n in range(20):
            eager.prepare_step(), eager.compute_step()
            loop.step()
            for name in ("y", "z"):
                np.testing.assert_array_equal(tb.to_numpy(graphed.state[name]),
                                              tb.to_numpy(eager.state[name]), err_msg=f"step {n}")
        assert loop.period == 6                    # lcm(3, 2) configurations, then replays only
        assert loop.replayed_launches == 19        # one launch per graphed step
    assert float(np.abs(tb.to_numpy(eager.state["y"]) - y0).max()) > 1.0


def test_moist_model_graphed_issues_the_eager_call_sequence():
    """Same kernels, same (canonical) pointers, same scalars, step after step, whether the moist
    model is stepped eagerly or through captured graphs."""
    from tasmania_b200.graphs import GraphedLoop
    from tasmania_b200.isentropic_moist import IsentropicMoistSUS

    nx, ny, nz = 17, 15, 8
    nsteps = 30
    grid, np_state = hp.moist_case(nx, ny, nz)
    traces = []
    with stubbed_library() as stub:
        FakeCapture.stub = stub
        for graphed in (False, True):
            grid.topography._fact = 0.0
            model = IsentropicMoistSUS(grid, np_state, timedelta(seconds=5), damp_depth=2)
            stub.trace = []
            if graphed:
                loop = GraphedLoop(model, capture_factory=FakeCapture)
                loop.run(nsteps)
                assert 2 <= loop.period <= 24, loop.period
                assert model.nstep == nsteps
                period = loop.period
            else:
                for _ in range(nsteps):
                    model.step()
            traces.append(canonical(stub.trace))
            assert model.state["time"] == model.init_time + nsteps * model.dt
            stub.trace = None
    assert len(traces[0]) == len(traces[1]) > 45 * nsteps  # 53 launches per step with the fused moist stage
    for n, (a, b) in enumerate(zip(*traces)):
        assert a == b, (n, a, b)
    print("buffer-rotation period of the moist SUS model:", period)


def test_dry_run_graphed_issues_the_eager_call_sequence():
    """configs[1]'s loop (dycore + diagnostics refresh): two buffer configurations (ping-pong)."""
    from tasmania_b200.graphs import GraphedLoop
    from tasmania_b200.grid import Grid, Topography, gaussian_profile, isentropic_state_from_brunt_vaisala
    from tasmania_b200.isentropic_dry import IsentropicDryRun

    nx, ny, nz, nsteps = 19, 17, 6, 9
    x, y = np.linspace(-176.0, 176.0, nx), np.linspace(-176.0, 176.0, ny)
    traces = []
    with stubbed_library() as stub:
        FakeCapture.stub = stub
        for graphed in (False, True):
            topo = Topography(gaussian_profile(x, y, 500.0, 50.0, 50.0), timedelta(seconds=30))
            grid = Grid((-176.0, 176.0), nx, (-176.0, 176.0), ny, (400.0, 280.0), nz, units_to_m=1e3,
                        topography=topo)
            np_state = isentropic_state_from_brunt_vaisala(grid, 22.5, 0.0, 0.015)
            run = IsentropicDryRun(grid, np_state, timedelta(seconds=5), damp_depth=2)
            stub.trace = []
            if graphed:
                loop = GraphedLoop(run, capture_factory=FakeCapture)
                loop.run(nsteps)
                assert loop.period == 2
                # the topography stops growing after 6 steps: from then on a step is ONE graph launch
                assert [n for n, _ in stub.trace].count("tb200_elementwise") == 2 * 6
            else:
                for _ in range(nsteps):
                    run.step()
            traces.append(canonical(stub.trace))
            assert run.state["time"] == run.init_time + nsteps * run.dt
            stub.trace = None
    assert len(traces[0]) >= 4 * nsteps
    assert traces[0] == traces[1]


def test_graphed_loop_refuses_a_step_that_never_cycles():
    """A model that allocates a new output array every step has no finite set of buffer
    configurations: the loop stops capturing at ``max_graphs`` instead of leaking graphs."""
    import pytest

    import tasmania_b200 as tb
    from tasmania_b200 import stencils
    from tasmania_b200.graphs import GraphedLoop

    class Leaky(Rotating):
        def compute_step(self):
            self.pool["a"] = tb.zeros(self.state["y"].shape)  # a fresh array each step
            super().compute_step()

    with stubbed_library() as stub:
        FakeCapture.stub = stub
        loop = GraphedLoop(Leaky(tb, stencils, np.ones((3, 2, 2))), eager_steps=0, max_graphs=4,
                           capture_factory=FakeCapture)
        keep = []  # hold the arrays so that the allocator cannot hand the same address out again
        with pytest.raises(tb.lib.B200Error, match="does not cycle"):
            for _ in range(10):
                loop.step()
                keep.append(dict(loop.model.pool))
        assert loop.period == 4
