# -*- coding: utf-8 -*-
"""Host-resident states through the device: the end-to-end entry point of the dry dynamical core.

A caller whose model state lives in host memory (the reference's numpy world) hands one state
per step to ``HostStreamedDryCore.step``: the stage inputs (s, su, sv, u, v, Montgomery
potential) are uploaded from pinned host buffers, one full RK step plus the diagnostics refresh
runs on the device, and the stepped fields (s, su, sv, u, v) are downloaded into pinned host
buffers.  The device side is double-buffered and the three activities run on three CUDA
streams, so the upload of step i+1 and the download of step i-1 overlap the computation of step
i (PCIe is full duplex): the sustained rate is max(upload, compute + download-of-previous) per
step instead of their sum.  ``bench.py`` reports this path as ``e2e``.
"""
from __future__ import annotations

from datetime import datetime

import torch

from tasmania_b200 import storage
from tasmania_b200.isentropic import MTG, S, SU, SV, U, V

P, EXN, H = ("air_pressure_on_interface_levels", "exner_function_on_interface_levels",
             "height_on_interface_levels")


def flat(arr):
    """The contiguous padded allocation behind a storage: host mirrors use the same layout, so a
    transfer is one plain cudaMemcpyAsync per field."""
    t = arr.t
    return t._base if t._base is not None else t


class HostStreamedDryCore:
    names_in = (S, SU, SV, U, V, MTG)
    names_out = (S, SU, U, SV, V)

    def __init__(self, dycore, diagnostics, pt, timestep, start_time=None):
        self.dyc, self.diag, self.pt, self.dt = dycore, diagnostics, pt, timestep
        shape = dycore.storage_shape
        dev = dycore.storage_options.device
        z = lambda: storage.zeros(shape, device=dev)  # noqa: E731
        self.sets = [{"in": {n: z() for n in self.names_in + (P, EXN, H)},
                      "out": {n: z() for n in self.names_out}} for _ in range(2)]
        self.h2d, self.d2h = torch.cuda.Stream(), torch.cuda.Stream()
        ev = lambda: [torch.cuda.Event(), torch.cuda.Event()]  # noqa: E731
        self.uploaded, self.in_free, self.computed, self.out_free = ev(), ev(), ev(), ev()
        self.time = start_time or datetime(2000, 1, 1)
        self.nstep = 0

    def host_buffers(self, names):
        """Pinned host buffers with the device layout, for ``step``."""
        ref = self.sets[0]["in"][S]
        return {n: torch.empty_like(flat(ref), device="cpu").pin_memory() for n in names}

    def step(self, host_in, host_out):
        """Enqueue upload -> RK step + diagnostics -> download; returns immediately."""
        b = self.nstep % 2
        dev = self.sets[b]
        main = torch.cuda.current_stream()
        self.h2d.wait_event(self.in_free[b])  # step i-2 has finished reading this input set
        with torch.cuda.stream(self.h2d):
            for n in self.names_in:
                flat(dev["in"][n]).copy_(host_in[n], non_blocking=True)
            self.uploaded[b].record()
        main.wait_event(self.uploaded[b])
        main.wait_event(self.out_free[b])     # download i-2 has finished reading this output set
        self.nstep += 1
        self.dyc.update_topography(self.nstep * self.dt)
        state = dict(dev["in"])
        state["time"] = self.time
        out = self.dyc(state, {}, self.dt, out_state=dev["out"])
        self.time = out["time"]
        self.diag.get_diagnostic_variables(out[S], self.pt, dev["in"][P], dev["in"][EXN],
                                           dev["in"][MTG], dev["in"][H])
        self.in_free[b].record(main)
        self.computed[b].record(main)
        self.d2h.wait_event(self.computed[b])
        with torch.cuda.stream(self.d2h):
            for n in self.names_out:
                host_out[n].copy_(flat(dev["out"][n]), non_blocking=True)
            self.out_free[b].record()

    def join(self):
        """Make the current stream wait for every enqueued download."""
        main = torch.cuda.current_stream()
        for e in self.out_free:
            main.wait_event(e)
