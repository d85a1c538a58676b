# -*- coding: utf-8 -*-
"""CUDA-graph replay of a model's time step.

On the small grids of BASELINE.json (configs[1]: 161x161x60, configs[2]: 256x256x60) a step is a
chain of 12 / ~120 kernels of 10-50 microseconds each, and the Python + ctypes launch path
(~45 microseconds per call) is slower than the device: measured on B200 (profiles/README.md, round
1d) the host needs 0.63 / 5.4 ms to enqueue a step whose kernels take a fraction of that.  The
reference has the same structure (one Python call per stencil).  The b200 answer is to capture the
step's kernel sequence once and replay it with one ``cudaGraphLaunch``.

A step is not one fixed sequence, though: output buffers rotate (the dynamical core ping-pongs its
state dicts, ``update_swap`` exchanges arrays between the state and every component's private
buffers), so consecutive steps run the same kernels on *different pointers*, with a period that
depends on the component list (2 for the dry core, 12 for the moist SUS model).  ``GraphedLoop``
therefore keys graphs by the buffer configuration: a step whose configuration has not been seen is
captured (and then replayed -- capture does not execute), a step whose configuration is known is
replayed and the model's dict bookkeeping is set to what the captured Python code left behind.
After one period every step is a single graph launch and no Python model code runs.

Protocol of a model (see tasmania_b200.isentropic_moist.IsentropicMoistSUS):

  prepare_step()      host-dependent work of a step, run eagerly before the graph: time level,
                      growth factor of the topography -> device (a tiny kernel only while it changes)
  compute_step()      everything else: a fixed kernel sequence given the buffer configuration
  finish_step()       host bookkeeping after the step (time stamp)
  buffer_dicts()      the dicts (name -> storage) whose arrays a step permutes
  set_buffer_dicts()  install a saved configuration

Results are bit-identical to eager stepping (tests/test_gpu_graphs.py): the graph holds the very
same kernel launches with the very same arguments.
"""
from __future__ import annotations

from tasmania_b200 import lib


def _ptr(arr):
    t = getattr(arr, "t", arr)
    return (t.data_ptr(), tuple(t.shape), tuple(t.stride()))


def buffer_key(dicts):
    """Hashable description of which array sits under which name in every dict."""
    return tuple(tuple(sorted((n, _ptr(v)) for n, v in d.items() if n != "time")) for d in dicts)


def snapshot(dicts):
    return [{n: v for n, v in d.items() if n != "time"} for d in dicts]


class _TorchCapture:
    """torch.cuda.CUDAGraph capture / replay (stream capture on torch's current stream, which is
    the stream every tasmania_b200 kernel is launched on)."""

    def __init__(self):
        import torch

        self._torch = torch
        self.graph = torch.cuda.CUDAGraph()

    def capture(self, fn):
        with self._torch.cuda.graph(self.graph):
            fn()

    def replay(self):
        self.graph.replay()


class GraphedLoop:
    def __init__(self, model, eager_steps=2, max_graphs=64, capture_factory=_TorchCapture):
        self.model = model
        self.eager_steps = int(eager_steps)  # first steps run eagerly: lazy allocations, module load
        self.max_graphs = int(max_graphs)
        self._factory = capture_factory
        self._graphs = {}  # buffer key -> (capture, snapshot after the step, launches in the graph)
        self.nsteps = 0
        self.replayed_launches = 0

    @property
    def period(self):
        """Number of distinct buffer configurations captured so far."""
        return len(self._graphs)

    def step(self):
        m = self.model
        self.nsteps += 1
        if self.nsteps <= self.eager_steps:
            m.prepare_step()
            m.compute_step()
            m.finish_step()
            return
        m.prepare_step()
        key = buffer_key(m.buffer_dicts())
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) >= self.max_graphs:
                raise lib.B200Error(
                    f"more than {self.max_graphs} buffer configurations: the step does not cycle")
            cap = self._factory()
            n0 = lib.launch_count()
            cap.capture(m.compute_step)  # records the launches; the Python bookkeeping advances
            entry = (cap, snapshot(m.buffer_dicts()), lib.launch_count() - n0)
            self._graphs[key] = entry
        else:
            # same configuration as at capture time: install what compute_step left behind then
            dicts = [dict(d) for d in entry[1]]
            old = m.buffer_dicts()
            for new, prev in zip(dicts, old):
                if "time" in prev:
                    new["time"] = prev["time"]
            m.set_buffer_dicts(dicts)
        entry[0].replay()
        self.replayed_launches += entry[2]
        m.finish_step()

    def run(self, nsteps):
        for _ in range(nsteps):
            self.step()
        return self.model.state
