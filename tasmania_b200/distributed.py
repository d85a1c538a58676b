# -*- coding: utf-8 -*-
"""2-D (x, y) domain decomposition of the isentropic dynamical core over the GPUs of one box
(SURVEY.md section 8e).  The reference has no distributed code at all; this is the part
``north_star`` adds.

Layout.  The global numerical grid ``NX x NY`` is cut into ``px x py`` blocks, one process (one
GPU) per block, x fastest in the rank order.  A rank's *local* grid is its owned block plus a
halo of ``HALO = nb + 1 = 4`` columns / rows on every side that has a neighbour (no halo on
physical boundaries, where the global relaxed boundary applies through the global relaxation
coefficients, cf. ``boundary.Relaxed(global_extent=..., offset=...)``).  The vertical is never
split: the column scans stay local.

Why ``nb + 1``.  A fused RK stage (isentropic_fused.cu) is: step s -> column scans -> momentum
step, and the momentum step reads the *new* Montgomery potential at i+-1 / j+-1.  Instead of a
second exchange in the middle of the stage, every rank also steps s (and scans the column) on
the first halo ring, which needs the stage inputs on ``nb`` more rings: one exchange of
``s, su, sv, u, v`` (width 4) per stage, none for the Montgomery potential.  Because that ring's
s-step reads the stage input on corner points, the exchange is done in two phases -- x faces
first, then y faces *including* the freshly received x halos -- which fills the corners without
diagonal messages.  Afterwards the velocity components on the two faces that separate owned
points from halo points are re-diagnosed locally (``tb200_velocity`` on a one-face box): the
fused kernel computed them from not-yet-exchanged momenta.

Every owned point sees bit-identical inputs and executes the same operations as in a single
-device run, so the decomposed result equals the single-device result **bitwise** -- the test
in tests/test_gpu_distributed.py (in-process sub-domains on one GPU) and the world_size-2
``gloo`` test of the exchange logic in tests/test_distributed_cpu.py.

Transport.  Default on GPUs (``TB200_HALO=p2p``): peer stores over NVLink -- one push launch
per phase packs the slabs of both sides straight into the neighbours' receive buffers (CUDA IPC
mappings) and raises their arrival counters, one pull launch waits for the own counters and
unpacks (``P2PHaloExchange``, csrc/halo.cu); no NCCL call and no host synchronisation on the
path.  ``TB200_HALO=nccl``: one pack kernel per side and phase gathers the slab of all fields
into one message (``tb200_halo_pack``); messages travel with ``torch.distributed``
point-to-point ops (grouped per phase); one unpack kernel per side scatters them.  There is no
global reduction on the path.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from datetime import datetime, timedelta
from typing import List, Sequence

import numpy as np
import torch

from tasmania_b200 import lib, storage

HALO = 4  # nb (fifth-order upwind: 3) + 1, see the module docstring


def process_grid(world: int):
    """1, 2, 4, 8 ranks -> 1x1, 2x1, 2x2, 4x2; otherwise the most
    square factorisation with px >= py."""
    py = int(np.floor(np.sqrt(world)))
    while world % py:
        py -= 1
    return world // py, py


def _split(n: int, parts: int):
    """Even split of n points over `parts` blocks (the first n % parts blocks get one more)."""
    base, rem = divmod(n, parts)
    edges = [0]
    for p in range(parts):
        edges.append(edges[-1] + base + (1 if p < rem else 0))
    return edges


@dataclass
class Side:
    """One neighbour of one phase: where the outgoing slab is read and the incoming one lands
    (local indices, (i0, j0) + (di, dj))."""

    name: str
    neighbour: int
    send_origin: tuple
    recv_origin: tuple
    extent: tuple


class Decomposition:
    """Index bookkeeping of the block decomposition (pure Python, shared by every backend)."""

    def __init__(self, nx_global: int, ny_global: int, px: int, py: int, halo: int = HALO):
        self.NX, self.NY, self.px, self.py, self.halo = nx_global, ny_global, px, py, halo
        self.xe, self.ye = _split(nx_global, px), _split(ny_global, py)
        for e in (self.xe, self.ye):
            if len(e) > 2:
                assert min(b - a for a, b in zip(e[:-1], e[1:])) >= 2 * halo, \
                    "blocks must be at least 2 * halo points wide"

    @property
    def world(self):
        return self.px * self.py

    def coords(self, rank):
        return rank % self.px, rank // self.px

    def rank_of(self, cx, cy):
        if 0 <= cx < self.px and 0 <= cy < self.py:
            return cy * self.px + cx
        return None

    def owned(self, rank):
        cx, cy = self.coords(rank)
        return self.xe[cx], self.xe[cx + 1], self.ye[cy], self.ye[cy + 1]

    def halos(self, rank):
        """(west, east, south, north) halo widths: `halo` towards a neighbour, 0 at a physical
        boundary."""
        cx, cy = self.coords(rank)
        h = self.halo
        return (h if cx > 0 else 0, h if cx < self.px - 1 else 0,
                h if cy > 0 else 0, h if cy < self.py - 1 else 0)

    def local(self, rank):
        """Global index range covered by the local grid: (i0, i1, j0, j1)."""
        i0, i1, j0, j1 = self.owned(rank)
        hw, he, hs, hn = self.halos(rank)
        return i0 - hw, i1 + he, j0 - hs, j1 + hn

    def local_shape(self, rank):
        i0, i1, j0, j1 = self.local(rank)
        return i1 - i0, j1 - j0

    def sides(self, rank, phase) -> List[Side]:
        """Phase 0: x faces over all local rows; phase 1: y faces over all local columns."""
        cx, cy = self.coords(rank)
        nxl, nyl = self.local_shape(rank)
        hw, he, hs, hn = self.halos(rank)
        h = self.halo
        out = []
        if phase == 0:
            if hw:
                out.append(Side("west", self.rank_of(cx - 1, cy), (hw, 0), (0, 0), (h, nyl)))
            if he:
                out.append(Side("east", self.rank_of(cx + 1, cy), (nxl - he - h, 0), (nxl - he, 0),
                                (h, nyl)))
        else:
            if hs:
                out.append(Side("south", self.rank_of(cx, cy - 1), (0, hs), (0, 0), (nxl, h)))
            if hn:
                out.append(Side("north", self.rank_of(cx, cy + 1), (0, nyl - hn - h), (0, nyl - hn),
                                (nxl, h)))
        return out

    def sides_single_phase(self, rank) -> List[Side]:
        """The same halo in ONE phase: x faces over the owned rows, y faces over the owned columns
        and the four corner blocks from the diagonal neighbours (what the two-phase plan routes
        through an x- and a y-neighbour).  Same final halo contents, half the synchronisations."""
        cx, cy = self.coords(rank)
        nxl, nyl = self.local_shape(rank)
        hw, he, hs, hn = self.halos(rank)
        h = self.halo
        ox0, ox1, oy0, oy1 = hw, nxl - he, hs, nyl - hn  # owned block, local indices
        out = []
        if hw:
            out.append(Side("west", self.rank_of(cx - 1, cy), (hw, oy0), (0, oy0), (h, oy1 - oy0)))
        if he:
            out.append(Side("east", self.rank_of(cx + 1, cy), (ox1 - h, oy0), (ox1, oy0), (h, oy1 - oy0)))
        if hs:
            out.append(Side("south", self.rank_of(cx, cy - 1), (ox0, hs), (ox0, 0), (ox1 - ox0, h)))
        if hn:
            out.append(Side("north", self.rank_of(cx, cy + 1), (ox0, oy1 - h), (ox0, oy1), (ox1 - ox0, h)))
        if hw and hs:
            out.append(Side("sw", self.rank_of(cx - 1, cy - 1), (hw, hs), (0, 0), (h, h)))
        if he and hs:
            out.append(Side("se", self.rank_of(cx + 1, cy - 1), (ox1 - h, hs), (ox1, 0), (h, h)))
        if hw and hn:
            out.append(Side("nw", self.rank_of(cx - 1, cy + 1), (hw, oy1 - h), (0, oy1), (h, h)))
        if he and hn:
            out.append(Side("ne", self.rank_of(cx + 1, cy + 1), (ox1 - h, oy1 - h), (ox1, oy1), (h, h)))
        return out

    def seam_faces(self, rank):
        """Staggered faces between an owned and a halo point, whose velocity must be
        re-diagnosed after the exchange: ([u face columns], [v face rows]), local indices."""
        nxl, nyl = self.local_shape(rank)
        hw, he, hs, hn = self.halos(rank)
        return ([hw] if hw else []) + ([nxl - he] if he else []), \
               ([hs] if hs else []) + ([nyl - hn] if hn else [])


# ------------------------------------------------------------------ pack / unpack
def _field_ptrs(fields):
    keep = [lib.as_field(f) for f in fields]
    import ctypes as C

    arr = (lib.FieldP * len(keep))(*[C.pointer(k) for k in keep])
    return arr, keep


def _pack(fields, buf, origin, extent, nz):
    if buf.is_cuda:
        arr, keep = _field_ptrs(fields)
        lib.check(lib.load().tb200_halo_pack(arr, len(fields), buf.data_ptr(),
                                             lib.int3((origin[0], origin[1], 0)),
                                             lib.int3((extent[0], extent[1], nz)),
                                             lib.current_stream()), "tb200_halo_pack")
        del keep
    else:  # host tensors: only the exchange-logic tests (gloo) come here
        i0, j0 = origin
        di, dj = extent
        view = buf.view(len(fields), nz, dj, di)
        for n, f in enumerate(fields):
            view[n].copy_(f.t[i0:i0 + di, j0:j0 + dj, :nz].permute(2, 1, 0))


def _unpack(fields, buf, origin, extent, nz):
    if buf.is_cuda:
        arr, keep = _field_ptrs(fields)
        lib.check(lib.load().tb200_halo_unpack(arr, len(fields), buf.data_ptr(),
                                               lib.int3((origin[0], origin[1], 0)),
                                               lib.int3((extent[0], extent[1], nz)),
                                               lib.current_stream()), "tb200_halo_unpack")
        del keep
    else:
        i0, j0 = origin
        di, dj = extent
        view = buf.view(len(fields), nz, dj, di)
        for n, f in enumerate(fields):
            f.t[i0:i0 + di, j0:j0 + dj, :nz].copy_(view[n].permute(2, 1, 0))


class HaloExchange:
    """The per-rank halo exchange of ``nfields`` fields: persistent message buffers, the
    two-phase plan and the three steps (pack, transfer, unpack) of each phase.  ``transfer``
    is either ``torch.distributed`` point-to-point (``exchange``) or done by the caller for
    in-process sub-domains (``exchange_in_process``)."""

    def __init__(self, decomp: Decomposition, rank: int, nz: int, nfields: int, device):
        self.decomp, self.rank, self.nz, self.nfields = decomp, rank, nz, nfields
        self.plan = [decomp.sides(rank, 0), decomp.sides(rank, 1)]
        self.send, self.recv = {}, {}
        for phase in self.plan:
            for s in phase:
                n = nfields * nz * s.extent[0] * s.extent[1]
                self.send[s.name] = torch.empty(n, dtype=torch.float64, device=device)
                self.recv[s.name] = torch.empty(n, dtype=torch.float64, device=device)
        self.bytes_per_exchange = 8 * sum(b.numel() for b in self.send.values())

    def message(self, buf, s, count):
        """The leading part of a persistent buffer that holds ``count`` fields of side ``s``
        (message layout [field][k][j][i]: fewer fields = a prefix)."""
        return buf[: count * self.nz * s.extent[0] * s.extent[1]]

    def pack(self, phase, fields):
        for s in self.plan[phase]:
            _pack(fields, self.message(self.send[s.name], s, len(fields)), s.send_origin, s.extent, self.nz)

    def unpack(self, phase, fields):
        for s in self.plan[phase]:
            _unpack(fields, self.message(self.recv[s.name], s, len(fields)), s.recv_origin, s.extent, self.nz)

    def transfer(self, phase, count=None):
        import torch.distributed as dist

        count = self.nfields if count is None else count
        ops = []
        for s in self.plan[phase]:
            ops.append(dist.P2POp(dist.isend, self.message(self.send[s.name], s, count), s.neighbour))
            ops.append(dist.P2POp(dist.irecv, self.message(self.recv[s.name], s, count), s.neighbour))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def exchange(self, fields: Sequence):
        """Two-phase exchange of up to ``nfields`` fields (every rank passes the same number)."""
        assert 1 <= len(fields) <= self.nfields
        for phase in (0, 1):
            self.pack(phase, fields)
            self.transfer(phase, len(fields))
            self.unpack(phase, fields)


_OPPOSITE = {"west": "east", "east": "west", "south": "north", "north": "south",
             "sw": "ne", "ne": "sw", "se": "nw", "nw": "se"}


def exchange_in_process(exchangers: Sequence[HaloExchange], fields_per_rank: Sequence[Sequence]):
    """All sub-domains live in this process (one GPU): a rank's outgoing message is unpacked
    straight from the neighbour's send buffer."""
    for phase in (0, 1):
        for ex, fields in zip(exchangers, fields_per_rank):
            ex.pack(phase, fields)
        for ex, fields in zip(exchangers, fields_per_rank):
            for s in ex.plan[phase]:
                src = exchangers[s.neighbour].send[_OPPOSITE[s.name]]
                _unpack(fields, ex.message(src, s, len(fields)), s.recv_origin, s.extent, ex.nz)


# ------------------------------------------------------------------ peer-to-peer transport
class P2PHaloExchange(HaloExchange):
    """The same two-phase plan over NVLink peer stores instead of NCCL messages (csrc/halo.cu,
    ``tb200_halo_push`` / ``tb200_halo_pull``): a phase is ONE push launch that packs the slabs of
    both sides straight into the neighbours' receive buffers and raises their arrival counters,
    and ONE pull launch that waits for the own counters and unpacks -- four launches per exchange,
    no send buffers, no host synchronisation, capturable in a CUDA graph.

    A rank owns one exportable allocation: per side two receive slots (exchange q uses slot
    q & 1), then one arrival counter per side (128 bytes apart), then the two per-phase channels
    (sequence numbers).  ``describe()`` is what the neighbours need (CUDA IPC handle + offsets);
    ``connect()`` takes the mapped base address and the description of every neighbour.
    """

    COUNTER_PITCH = 128

    def __init__(self, decomp: Decomposition, rank: int, nz: int, nfields: int, device=None, phases=None):
        import ctypes as C

        self.decomp, self.rank, self.nz, self.nfields = decomp, rank, nz, nfields
        # one phase (faces + corner blocks from the diagonal neighbours: one push, one pull, one
        # wait per exchange; default) or the two phases of the message-based exchange
        # (TB200_HALO_PHASES=2)
        self.phases = int(phases or os.environ.get("TB200_HALO_PHASES", "1"))
        assert self.phases in (1, 2)
        self.plan = ([decomp.sides_single_phase(rank), []] if self.phases == 1
                     else [decomp.sides(rank, 0), decomp.sides(rank, 1)])
        self.send, self.recv = {}, {}
        self.layout, off = {}, 0
        for phase in self.plan:
            for s in phase:
                slot = nfields * nz * s.extent[0] * s.extent[1]
                self.layout[s.name] = {"buffer": off, "slot_doubles": slot}
                off += 2 * slot * 8
        self.bytes_per_exchange = off // 2
        off = (off + 127) // 128 * 128
        for name in self.layout:
            self.layout[name]["counter"] = off
            off += self.COUNTER_PITCH
        self.channel_offset = [off, off + lib.P2P_CHANNEL_BYTES]
        off += 2 * lib.P2P_CHANNEL_BYTES
        self.nbytes = off
        base = C.c_void_p()
        lib.check(lib.load().tb200_p2p_alloc(self.nbytes, C.byref(base)), "tb200_p2p_alloc")
        self.base = int(base.value)
        self.sides_c = [None, None]
        self._imported = []

    def describe(self):
        """What a neighbour needs to push into this rank's buffers."""
        import ctypes as C

        handle = C.create_string_buffer(lib.P2P_HANDLE_BYTES)
        lib.check(lib.load().tb200_p2p_export(C.c_void_p(self.base), handle), "tb200_p2p_export")
        return {"rank": self.rank, "handle": handle.raw, "layout": self.layout}

    def connect(self, peers):
        """``peers``: neighbour rank -> (base address of its allocation as seen from THIS
        process, its ``layout``)."""
        for phase, sides in enumerate(self.plan):
            arr = (lib.HaloSide * max(len(sides), 1))()
            for m, s in enumerate(sides):
                pbase, playout = peers[s.neighbour]
                theirs, mine = playout[_OPPOSITE[s.name]], self.layout[s.name]
                assert theirs["slot_doubles"] == mine["slot_doubles"], "asymmetric halo plan"
                h = arr[m]
                h.remote_buffer = pbase + theirs["buffer"]
                h.remote_counter = pbase + theirs["counter"]
                h.local_buffer = self.base + mine["buffer"]
                h.local_counter = self.base + mine["counter"]
                h.slot_doubles = mine["slot_doubles"]
                h.send_origin[:] = list(s.send_origin)
                h.recv_origin[:] = list(s.recv_origin)
                h.extent[:] = list(s.extent)
            self.sides_c[phase] = arr

    def connect_in_process(self, exchangers):
        """All sub-domains in this process (one GPU): the neighbours' buffers are plain addresses."""
        self.connect({ex.rank: (ex.base, ex.layout) for ex in exchangers})

    def connect_ipc(self, group=None):
        """One process per GPU: descriptions travel through ``torch.distributed`` (any backend),
        neighbours' allocations are mapped with CUDA IPC (peer access over NVLink)."""
        import ctypes as C

        import torch.distributed as dist

        mine = self.describe()
        everyone = [None] * dist.get_world_size(group)
        dist.all_gather_object(everyone, mine, group=group)
        peers, failure = {}, None
        for phase in self.plan:
            for s in phase:
                if s.neighbour in peers or failure is not None:
                    continue
                d = everyone[s.neighbour]
                ptr = C.c_void_p()
                try:
                    lib.check(lib.load().tb200_p2p_import(d["handle"], C.byref(ptr)), "tb200_p2p_import")
                except lib.B200Error as exc:  # no peer access between the two devices
                    failure = exc
                    continue
                self._imported.append(int(ptr.value))
                peers[s.neighbour] = (int(ptr.value), d["layout"])
        # every rank learns whether every rank could map its neighbours (a collective, so that no
        # rank is left waiting at the barrier below), then nobody pushes before every mapping exists
        verdicts = [None] * dist.get_world_size(group)
        dist.all_gather_object(verdicts, None if failure is None else str(failure), group=group)
        bad = [v for v in verdicts if v is not None]
        if bad:
            raise lib.B200Error(f"peer-store halo transport unavailable: {bad[0]}")
        self.connect(peers)
        dist.barrier(group)

    def _call(self, fn, what, phase, fields):
        sides = self.plan[phase]
        if not sides:
            return
        arr, keep = _field_ptrs(fields)
        lib.check(fn(arr, len(fields), self.sides_c[phase], len(sides),
                     self.base + self.channel_offset[phase], 0, self.nz, lib.current_stream()), what)
        del keep

    def push(self, phase, fields):
        self._call(lib.load().tb200_halo_push, "tb200_halo_push", phase, fields)

    def pull(self, phase, fields):
        self._call(lib.load().tb200_halo_pull, "tb200_halo_pull", phase, fields)

    def exchange(self, fields: Sequence):
        assert 1 <= len(fields) <= self.nfields
        for phase in (0, 1):
            self.push(phase, fields)
            self.pull(phase, fields)

    def check(self):
        """Raise if a pull ever gave up waiting for its peer (device synchronised by the copy)."""
        import ctypes as C

        for phase in (0, 1):
            err = C.c_int(0)
            lib.check(lib.load().tb200_p2p_channel_error(self.base + self.channel_offset[phase], C.byref(err)),
                      "tb200_p2p_channel_error")
            if err.value:
                raise lib.B200Error(f"halo exchange (rank {self.rank}, phase {phase}): a neighbour's slab never "
                                    "arrived -- the results are invalid")

    def close(self):
        for ptr in self._imported:
            lib.load().tb200_p2p_release(ptr)
        self._imported = []
        if self.base:
            lib.load().tb200_p2p_free(self.base)
            self.base = 0


def exchange_in_process_p2p(exchangers: Sequence[P2PHaloExchange], fields_per_rank: Sequence[Sequence]):
    """All sub-domains in one process on one GPU: every push of a phase is enqueued before the
    first pull, so no pull ever waits (kernels of one GPU must not wait on one another)."""
    for phase in (0, 1):
        for ex, fields in zip(exchangers, fields_per_rank):
            ex.push(phase, fields)
        for ex, fields in zip(exchangers, fields_per_rank):
            ex.pull(phase, fields)


def halo_transport():
    """``TB200_HALO=p2p`` (default on GPUs) | ``nccl``: peer stores over NVLink or
    ``torch.distributed`` point-to-point messages."""
    v = os.environ.get("TB200_HALO", "p2p").lower()
    if v not in ("p2p", "nccl"):
        raise ValueError(f"TB200_HALO must be p2p or nccl, got {v!r}")
    return v


# ------------------------------------------------------------------ the decomposed dry core
P, EXN, H = ("air_pressure_on_interface_levels", "exner_function_on_interface_levels",
             "height_on_interface_levels")


class SubdomainDryCore:
    """One rank's share of the dry mountain-flow problem (BASELINE configs 2 / 5): local grid,
    local window of the global topography / relaxation coefficients / initial state, the fused
    dynamical core on it, and the per-stage halo exchange + seam-face velocity fix-up."""

    def __init__(self, decomp: Decomposition, rank: int, nz: int, *, domain_x=(-176.0, 176.0),
                 domain_y=(-176.0, 176.0), nb=3, nr=6, flux="fifth_order_upwind",
                 scheme="rk3ws_si", damp_depth=15, damp_max=5e-4, dt_seconds=5.0,
                 mountain=(500.0, 50.0, 50.0), topo_seconds=1800.0, device=None, transport="nccl"):
        import tasmania_b200 as tb
        from tasmania_b200.boundary import Relaxed
        from tasmania_b200.grid import (Grid, Topography, gaussian_profile,
                                        isentropic_state_from_brunt_vaisala)
        from tasmania_b200.isentropic import (MTG, S, SU, SV, U, V, IsentropicDiagnostics,
                                              IsentropicDynamicalCore)

        assert decomp.halo >= nb + 1 or decomp.world == 1
        self.decomp, self.rank, self.nz = decomp, rank, nz
        self.names = (S, SU, SV, U, V, MTG)
        self.out_names = (S, SU, U, SV, V)
        self.S, self.SU, self.SV, self.U, self.V, self.MTG = S, SU, SV, U, V, MTG
        NX, NY = decomp.NX, decomp.NY
        gi0, gi1, gj0, gj1 = decomp.local(rank)
        nx, ny = gi1 - gi0, gj1 - gj0
        self.nx, self.ny = nx, ny
        self.owned_points = (decomp.owned(rank)[1] - decomp.owned(rank)[0]) * \
                            (decomp.owned(rank)[3] - decomp.owned(rank)[2]) * nz
        # global axes (native units: km), local windows of them
        xg = np.linspace(domain_x[0], domain_x[1], NX)
        yg = np.linspace(domain_y[0], domain_y[1], NY)
        full = Grid(domain_x, NX, domain_y, NY, (400.0, 280.0), nz, units_to_m=1e3)
        steady = gaussian_profile(xg[gi0:gi1], yg[gj0:gj1], mountain[0], mountain[1], mountain[2],
                                  center_x=0.5 * (xg[0] + xg[-1]), center_y=0.5 * (yg[0] + yg[-1]))
        grid = Grid((xg[gi0], xg[gi1 - 1]), nx, (yg[gj0], yg[gj1 - 1]), ny, (400.0, 280.0), nz,
                    units_to_m=1e3, x=xg[gi0:gi1], y=yg[gj0:gj1],
                    topography=Topography(steady, timedelta(seconds=topo_seconds)))
        # grid spacings are those of the GLOBAL grid, bit for bit
        grid.dx_native, grid.dy_native, grid.dx, grid.dy = full.dx_native, full.dy_native, full.dx, full.dy
        self.grid = grid
        # horizontally uniform initial state: one column, broadcast
        small = Grid(domain_x, 3, domain_y, 3, (400.0, 280.0), nz, units_to_m=1e3)
        col = isentropic_state_from_brunt_vaisala(small, 22.5, 0.0, 0.015)
        self.pt = float(col[P][0, 0, 0])
        so = tb.StorageOptions(device=device)
        state = {}
        for name, a in col.items():
            arr = np.zeros((nx + 1, ny + 1, nz + 1))
            # staggered extents in GLOBAL terms: the extra face exists on the last block only
            mi = min(nx + 1, (NX + 1 if "at_u_locations" in name else NX) - gi0)
            mj = min(ny + 1, (NY + 1 if "at_v_locations" in name else NY) - gj0)
            arr[:mi, :mj, :] = a[1, 1, :][None, None, :]
            state[name] = tb.as_storage(arr, device=device)
        state["time"] = datetime(2000, 1, 1)
        self.state = state
        self.dt = timedelta(seconds=dt_seconds)
        self.hb = Relaxed(nx, ny, nz, nb, nr=nr, storage_options=so, global_extent=(NX, NY),
                          offset=(gi0, gj0))
        self.hb.reference_state = state
        self.dyc = IsentropicDynamicalCore(
            grid, self.hb, time_integration_scheme=scheme, horizontal_flux_scheme=flux,
            time_integration_properties={"pt": self.pt, "eps": 0.5}, damp=True,
            damp_depth=damp_depth, damp_max=damp_max, storage_options=so)
        assert self.dyc._fused
        self.diag = IsentropicDiagnostics(grid, storage_options=so)
        self.spare = {n: tb.zeros(self.dyc.storage_shape, device=device) for n in self.out_names}
        dev = state[S].t.device
        self.transport = transport if dev.type == "cuda" else "nccl"
        if self.transport == "p2p":
            self.halo = P2PHaloExchange(decomp, rank, nz, 5, dev)  # connected by the owner of the run
        else:
            self.halo = HaloExchange(decomp, rank, nz, 5, dev)
        self.u_faces, self.v_faces = decomp.seam_faces(rank)
        self.nstep = 0

    # ---- pieces of a step (driven stage by stage so that sub-domains can interleave)
    def velocities_written(self, stage):
        """Does the stage kernel write u, v?  (With lazy velocities no stage does: a successor
        re-diagnoses them from the exchanged s, su, sv, and the step's final state gets them from
        one pass over the exchanged fields.)"""
        return not self.dyc.lazy_velocities

    def exchange_fields(self, out, stage=None):
        names = (self.S, self.SU, self.SV, self.U, self.V) if self.velocities_written(stage) \
            else (self.S, self.SU, self.SV)
        return [out[n] for n in names]

    def fix_seam_velocities(self, out, stage=None):
        """After the exchange.  Lazy velocities: u, v of the step's final state over the whole
        local grid, halos included, from the exchanged s, su, sv.  Otherwise: u on the faces
        between owned and halo columns, v likewise in y, which the fused kernel computed from
        not-yet-exchanged momenta (dwarfs/diagnostics.py:L219-L272, same formula)."""
        if self.dyc.lazy_velocities:
            if stage is None or stage == self.dyc.stages - 1:
                self.dyc.diagnose_velocities(out)
            return
        vc = self.dyc._velocity_components
        for i in self.u_faces:
            vc._stencil_diagnosing_velocity_x(in_d=out[self.S], in_du=out[self.SU], out_u=out[self.U],
                                              origin=(i, 0, 0), domain=(1, self.ny, self.nz))
        for j in self.v_faces:
            vc._stencil_diagnosing_velocity_y(in_d=out[self.S], in_dv=out[self.SV], out_v=out[self.V],
                                              origin=(0, j, 0), domain=(self.nx, 1, self.nz))

    def begin_step(self):
        self.nstep += 1
        self.dyc.update_topography(self.nstep * self.dt)
        return self.dyc.stages_iter(self.state, {}, self.dt, out_state=self.spare)

    def end_step(self, out):
        new = {n: out[n] for n in self.out_names}
        new["time"] = out["time"]
        for n in (P, EXN, H, self.MTG):
            new[n] = self.state[n]
        self.spare = {n: self.state[n] for n in self.out_names}
        self.diag.get_diagnostic_variables(new[self.S], self.pt, new[P], new[EXN], new[self.MTG], new[H])
        self.state = new

    def owned_numpy(self, name):
        """The owned block of a field on the host (for gathering / comparing)."""
        hw, he, hs, hn = self.decomp.halos(self.rank)
        a = storage.to_numpy(self.state[name])
        return a[hw:self.nx - he, hs:self.ny - hn, :self.nz]


class Overlap:
    """Runs the halo exchange of a stage on a side stream, under the interior blocks of the
    momentum kernel (IsentropicDynamicalCore.stage_array_call drives it):

        main stream   A, B, MV(rim blocks) | MV(interior blocks) ............ | next stage
        side stream                        | pack, send/recv, unpack, seam fix-up |

    ``rim`` = local columns / rows next to each edge with a neighbour that must be final before
    the exchange starts: the halo itself plus the ``halo`` owned points the neighbour receives.
    """

    def __init__(self, sub):
        self.sub = sub
        hw, he, hs, hn = sub.decomp.halos(sub.rank)
        h = sub.decomp.halo
        self.rim = tuple(x + h if x else 0 for x in (hw, he, hs, hn))
        self.side = torch.cuda.Stream()
        self.rim_done = torch.cuda.Event()
        self.exchanged = torch.cuda.Event()
        self.events = None  # optional (start, end) timing events of the exchanges

    def after_rim(self, stage, out):
        main = torch.cuda.current_stream()
        self.rim_done.record(main)
        self.side.wait_event(self.rim_done)
        with torch.cuda.stream(self.side):
            if self.events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            self.sub.halo.exchange(self.sub.exchange_fields(out, stage))
            if self.events is not None:
                e1.record()
                self.events.append((e0, e1))
            if not self.sub.dyc.lazy_velocities:  # seam faces: rim data only
                self.sub.fix_seam_velocities(out, stage)
            self.exchanged.record(self.side)

    def after_interior(self, stage, out):
        # whatever comes next on the main stream reads the exchanged halos
        torch.cuda.current_stream().wait_event(self.exchanged)
        if self.sub.dyc.lazy_velocities:  # u, v of the step's final state: needs the interior too
            self.sub.fix_seam_velocities(out, stage)


class DecomposedDryRun:
    """The timed loop of ``bench.py`` on N GPUs: one ``SubdomainDryCore`` per process,
    ``torch.distributed`` (NCCL) halo exchange after every RK stage.  Weak scaling: every rank
    owns ``nx x ny x nz`` points; the global domain grows with the process grid so that the
    grid spacing -- and with it the physics per point -- stays the same (2.2 km, config 2)."""

    def __init__(self, nx, ny, nz, rank, world, device=None, overlap=False, transport=None, **kwargs):
        px, py = process_grid(world)
        self.decomposition = f"{px}x{py}"
        self.decomp = Decomposition(nx * px, ny * py, px, py)
        # config 2's grid spacing (2.2 km) on the global grid: dt = 5 s stays stable at any size
        hx, hy = 1.1 * (nx * px - 1), 1.1 * (ny * py - 1)
        self.domain_x, self.domain_y = (-hx, hx), (-hy, hy)
        self.transport = (transport or halo_transport()) if world > 1 else "nccl"
        self.sub = SubdomainDryCore(self.decomp, rank, nz, domain_x=self.domain_x, domain_y=self.domain_y,
                                    device=device, transport=self.transport, **kwargs)
        if self.transport == "p2p":
            try:
                self.sub.halo.connect_ipc()
            except lib.B200Error as exc:
                # raised on EVERY rank together (connect_ipc): fall back to the message transport
                import warnings

                warnings.warn(f"{exc}; falling back to TB200_HALO=nccl")
                self.sub.halo.close()
                self.transport = self.sub.transport = "nccl"
                self.sub.halo = HaloExchange(self.decomp, rank, nz, 5, self.sub.state[self.sub.S].t.device)
        self.nx, self.ny, self.nz = nx, ny, nz
        self.names, self.out_names = self.sub.names, self.sub.out_names
        self.dyc = self.sub.dyc
        self.exchange_events = None  # set to [] to time the exchanges with CUDA events
        self.overlap = Overlap(self.sub) if (overlap and world > 1) else None
        if self.overlap is not None:
            self.dyc.overlap = self.overlap
        else:
            self.dyc.after_stage = self._after_stage

    def _after_stage(self, stage, out):
        if os.environ.get("TB200_SKIP_EXCHANGE"):  # timing experiments only: results are wrong
            return
        if self.exchange_events is not None:  # optional device timing of the exchange
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self.sub.halo.exchange(self.sub.exchange_fields(out, stage))
        if self.exchange_events is not None:
            e1.record()
            self.exchange_events.append((e0, e1))
        self.sub.fix_seam_velocities(out, stage)

    def exchange_ms(self):
        """Mean device time of one halo exchange (after a synchronize): packing, transfer, waiting for
        the neighbours' slabs, unpacking."""
        ev = (self.overlap.events if self.overlap is not None else self.exchange_events) or []
        return float(np.mean([a.elapsed_time(b) for a, b in ev])) if ev else None

    @property
    def state(self):
        return self.sub.state

    def step(self):
        out = None
        for _, out in self.sub.begin_step():
            pass
        self.sub.end_step(out)


class InProcessDecomposedRun:
    """All sub-domains of a decomposition in ONE process on one device -- the bitwise-parity
    harness of the decomposition (and a way to run a decomposed case without several GPUs)."""

    def __init__(self, nx_global, ny_global, nz, px, py, device=None, transport="nccl", **kwargs):
        self.decomp = Decomposition(nx_global, ny_global, px, py)
        self.subs = [SubdomainDryCore(self.decomp, r, nz, device=device, transport=transport, **kwargs)
                     for r in range(px * py)]
        self.transport = self.subs[0].transport
        if self.transport == "p2p":
            for s in self.subs:
                s.halo.connect_in_process([t.halo for t in self.subs])

    def step(self):
        iters = [s.begin_step() for s in self.subs]
        outs = [None] * len(self.subs)
        for stage in range(self.subs[0].dyc.stages):
            for r, it in enumerate(iters):
                outs[r] = next(it)[1]
            (exchange_in_process_p2p if self.transport == "p2p" else exchange_in_process)(
                [s.halo for s in self.subs], [s.exchange_fields(o, stage) for s, o in zip(self.subs, outs)])
            for s, o in zip(self.subs, outs):
                s.fix_seam_velocities(o, stage)
        for it in iters:  # exhaust the generators (sets the time label)
            for _ in it:
                pass
        for s, o in zip(self.subs, outs):
            s.end_step(o)

    def gather(self, name):
        """Assemble the global field from the owned blocks."""
        d = self.decomp
        out = np.zeros((d.NX, d.NY, self.subs[0].nz))
        for s in self.subs:
            i0, i1, j0, j1 = d.owned(s.rank)
            out[i0:i1, j0:j1, :] = s.owned_numpy(name)
        return out
