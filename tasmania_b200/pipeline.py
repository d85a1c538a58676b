# -*- coding: utf-8 -*-
"""Host-resident states through the device: the end-to-end entry point of the dry dynamical core.

A caller whose model state lives in host memory (the reference's numpy world) hands one state
per step to ``HostStreamedDryCore.step``: the PROGNOSTIC fields (s, su, sv) are uploaded from
pinned host buffers, one full RK step plus the diagnostics refresh runs on the device, and the
stepped prognostic fields are downloaded into pinned host buffers.  Everything else the step
needs is a diagnosis of those three and is made on the device, where it costs a fraction of its
PCIe transfer: the Montgomery potential by the column-scan kernel (40 B/pt of HBM traffic against
8 B/pt over PCIe, i.e. 0.5 ms against 10 ms at 1024 x 1024 x 64), the advecting velocities inside
the stage kernels (``derive_uv_in``: u = (su[i-1] + su[i]) / (s[i-1] + s[i]), the formula every
stage of the reference ends with).  ``prognostic_only=False`` restores the round-1 behaviour
(upload s, su, sv, u, v, mtg; download s, su, u, sv, v), for callers whose velocities are NOT that
diagnosis (an initial state given in terms of u, v).  The device side is double-buffered and the three activities run on three CUDA
streams, so the upload of step i+1 and the download of step i-1 overlap the computation of step
i (PCIe is full duplex): the sustained rate is max(upload, compute + download-of-previous) per
step instead of their sum.  ``bench.py`` reports this path as ``e2e``.
"""
from __future__ import annotations

import os
from datetime import datetime

import torch

from tasmania_b200 import storage
from tasmania_b200.isentropic import MTG, S, SU, SV, U, V

P, EXN, H = ("air_pressure_on_interface_levels", "exner_function_on_interface_levels",
             "height_on_interface_levels")


def gpu_numa_node(device_index=None):
    """NUMA node of the host memory closest to a GPU (sysfs ``numa_node`` of its PCI function), or
    None when it cannot be told (single-node host, virtualised PCI topology, no sysfs)."""
    cands = []
    try:
        idx = torch.cuda.current_device() if device_index is None else device_index
        pr = torch.cuda.get_device_properties(idx)
        dom, bus, dev = (getattr(pr, "pci_domain_id", None), getattr(pr, "pci_bus_id", None),
                         getattr(pr, "pci_device_id", None))
        if bus is not None and dev is not None:
            cands.append(f"{(dom or 0):04x}:{bus:02x}:{dev:02x}.0")
    except Exception:  # noqa: BLE001  (best effort: placement is an optimisation)
        pass
    try:
        import pynvml

        pynvml.nvmlInit()
        idx = torch.cuda.current_device() if device_index is None else device_index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis and all(v.strip().isdigit() for v in vis.split(",")):
            idx = int(vis.split(",")[idx])
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        cands += [bus, bus[4:] if len(bus) > 12 else bus]  # NVML pads the PCI domain to 8 digits
    except Exception:  # noqa: BLE001
        pass
    for cand in cands:
        try:
            with open(f"/sys/bus/pci/devices/{cand}/numa_node") as f:
                node = int(f.read().strip())
            if node >= 0:
                return node
        except (OSError, ValueError):
            continue
    return None


class prefer_numa_node:
    """``with prefer_numa_node(n):`` -- pages first touched inside (pinned allocations included) are
    taken from NUMA node ``n`` when it has room (set_mempolicy(MPOL_PREFERRED), x86-64 / aarch64
    Linux; a no-op anywhere else or when ``n`` is None).  ``applied`` tells whether the policy was set."""

    _SYS = {"x86_64": 238, "aarch64": 237}

    def __init__(self, node):
        self.node, self.applied = node, False

    def _set(self, mode, node):
        import ctypes
        import platform

        nr = self._SYS.get(platform.machine())
        if nr is None:
            return False
        libc = ctypes.CDLL(None, use_errno=True)
        if node is None:
            return libc.syscall(nr, 0, None, 0) == 0  # MPOL_DEFAULT
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        return libc.syscall(nr, mode, ctypes.byref(mask), 16 * 64 + 1) == 0

    def __enter__(self):
        if self.node is not None:
            try:
                self.applied = bool(self._set(1, self.node))  # MPOL_PREFERRED
            except Exception:  # noqa: BLE001
                self.applied = False
        return self

    def __exit__(self, *exc):
        if self.applied:
            try:
                self._set(0, None)
            except Exception:  # noqa: BLE001
                pass
        return False


def flat(arr):
    """The contiguous padded allocation behind a storage: host mirrors use the same layout, so a
    transfer is one plain cudaMemcpyAsync per field."""
    t = arr.t
    return t._base if t._base is not None else t


class HostStreamedDryCore:
    def __init__(self, dycore, diagnostics, pt, timestep, start_time=None, prognostic_only=True):
        self.dyc, self.diag, self.pt, self.dt = dycore, diagnostics, pt, timestep
        self.prognostic_only = bool(prognostic_only)
        if self.prognostic_only and not (dycore._fused and dycore.lazy_velocities):
            raise ValueError("prognostic_only needs the fused dynamical core with lazy velocities")
        self.names_in = (S, SU, SV) if self.prognostic_only else (S, SU, SV, U, V, MTG)
        self.names_out = (S, SU, SV) if self.prognostic_only else (S, SU, U, SV, V)
        self.device_names_in = (S, SU, SV, U, V, MTG)   # what a device-side state dict holds
        self.device_names_out = (S, SU, U, SV, V)
        shape = dycore.storage_shape
        dev = dycore.storage_options.device
        z = lambda: storage.zeros(shape, device=dev)  # noqa: E731
        self.sets = [{"in": {n: z() for n in self.device_names_in + (P, EXN, H)},
                      "out": {n: z() for n in self.device_names_out}} for _ in range(2)]
        self.h2d, self.d2h = torch.cuda.Stream(), torch.cuda.Stream()
        ev = lambda: [torch.cuda.Event(), torch.cuda.Event()]  # noqa: E731
        self.uploaded, self.in_free, self.computed, self.out_free = ev(), ev(), ev(), ev()
        self.time = start_time or datetime(2000, 1, 1)
        self.nstep = 0

    def host_buffers(self, names):
        """Pinned host buffers with the device layout, for ``step``; placed on the NUMA node next
        to this process's GPU when that can be told (eight ranks streaming 3.3 GB per step each
        through one node's memory controllers is what limits the 8-GPU end-to-end rate otherwise).
        ``self.numa_node`` = the node used, or None."""
        ref = self.sets[0]["in"][S]
        node = gpu_numa_node(flat(ref).device.index)
        with prefer_numa_node(node) as pol:
            out = {}
            for n in names:
                t = torch.empty_like(flat(ref), device="cpu").pin_memory()
                t.zero_()  # first touch under the policy
                out[n] = t
        self.numa_node = node if pol.applied else None
        return out

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
        state = dict(dev["in"])
        state["time"] = self.time
        if self.prognostic_only:
            # Montgomery potential of the uploaded s (and p, exn, h with it) over the topography
            # of the uploaded state's own time, i.e. BEFORE the topography moves on to this
            # step's; the velocities of stage 0 are diagnosed inside the kernels (u, v of this
            # set are never read)
            self.diag.get_diagnostic_variables(state[S], self.pt, state[P], state[EXN], state[MTG], state[H])
        self.nstep += 1
        self.dyc.update_topography(self.nstep * self.dt)
        if self.prognostic_only:
            saved = self.dyc.derive_stage0_velocities
            self.dyc.derive_stage0_velocities = True
            try:
                out = self.dyc(state, {}, self.dt, out_state=dev["out"])
            finally:
                self.dyc.derive_stage0_velocities = saved
        else:
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
