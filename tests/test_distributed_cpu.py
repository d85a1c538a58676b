# -*- coding: utf-8 -*-
"""Host-side logic of the 2-D domain decomposition, on CPU tensors:
the index plan of ``Decomposition``, the two-phase halo exchange in process and across two
``gloo`` ranks, and the windowed relaxation coefficients.  The kernels themselves are covered
by the ``-m gpu`` tests (bitwise equality of a decomposed and a single-device run)."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import boundary as ob
from tasmania_b200 import storage
from tasmania_b200.boundary import relaxed_gamma_window
from tasmania_b200.distributed import (Decomposition, HaloExchange, exchange_in_process,
                                       process_grid)

NF = 5


def _global_field(n, NX, NY, nz):
    i, j, k = np.meshgrid(np.arange(NX + 1), np.arange(NY + 1), np.arange(nz + 1), indexing="ij")
    return 1000.0 * n + i + 0.001 * j + 1e-6 * k


def _local_fields(d, rank, nz, fill_halo):
    """Local storages holding the global function on the owned block (and on the halos too if
    `fill_halo`), -1 elsewhere."""
    gi0, gi1, gj0, gj1 = d.local(rank)
    hw, he, hs, hn = d.halos(rank)
    nx, ny = gi1 - gi0, gj1 - gj0
    out = []
    for n in range(NF):
        g = _global_field(n, d.NX, d.NY, nz)
        a = np.full((nx + 1, ny + 1, nz + 1), -1.0)
        if fill_halo:
            a[:nx, :ny, :nz] = g[gi0:gi1, gj0:gj1, :nz]
        else:
            a[hw:nx - he, hs:ny - hn, :nz] = g[gi0 + hw:gi1 - he, gj0 + hs:gj1 - hn, :nz]
        out.append(storage.as_storage(a, device="cpu"))
    return out


def _check(d, rank, fields, nz):
    want = _local_fields(d, rank, nz, fill_halo=True)
    for f, w in zip(fields, want):
        np.testing.assert_array_equal(f.to_numpy(), w.to_numpy())


def test_process_grid():
    assert [process_grid(n) for n in (1, 2, 4, 8)] == [(1, 1), (2, 1), (2, 2), (4, 2)]


@pytest.mark.parametrize("NX,NY,px,py", [(16, 9, 2, 1), (9, 16, 1, 2), (19, 17, 2, 2),
                                          (40, 23, 4, 2), (27, 30, 3, 3)])
def test_exchange_in_process_fills_all_halos(NX, NY, px, py):
    nz = 3
    d = Decomposition(NX, NY, px, py)
    # blocks tile the grid exactly
    cover = np.zeros((NX, NY), dtype=int)
    for r in range(d.world):
        i0, i1, j0, j1 = d.owned(r)
        cover[i0:i1, j0:j1] += 1
    assert (cover == 1).all()
    fields = [_local_fields(d, r, nz, fill_halo=False) for r in range(d.world)]
    ex = [HaloExchange(d, r, nz, NF, "cpu") for r in range(d.world)]
    exchange_in_process(ex, fields)
    for r in range(d.world):
        _check(d, r, fields[r], nz)  # corners included: x phase first, then y over the x halos


def test_seam_faces_and_sides():
    d = Decomposition(32, 24, 2, 2)
    # rank 0 = south-west block: halos towards east and north only
    assert d.halos(0) == (0, 4, 0, 4)
    assert d.local(0) == (0, 20, 0, 16)
    assert d.seam_faces(0) == ([16], [12])
    assert d.seam_faces(3) == ([4], [4])
    east = d.sides(0, 0)[0]
    west = d.sides(1, 0)[0]
    # what rank 0 sends east is exactly the window rank 1 receives from the west
    assert east.neighbour == 1 and west.neighbour == 0 and east.extent == west.extent
    g0, g1 = d.local(0), d.local(1)
    assert g0[0] + east.send_origin[0] == g1[0] + west.recv_origin[0]
    assert g1[0] + west.send_origin[0] == g0[0] + east.recv_origin[0]


@pytest.mark.parametrize("nx,ny,nb,nr", [(20, 17, 3, 6), (33, 29, 2, 8), (16, 16, 1, 1)])
def test_windowed_gamma_equals_reference_construction(nx, ny, nb, nr):
    g = ob.relaxed_gamma(nx, ny, 4, nb, nr)
    g = g[:, :, 0] if g.ndim == 3 else g
    rel = np.array([1.0] + [1.0 - np.tanh(0.5 * m) for m in range(1, 8)])[:nr]
    rel[:nb] = 1.0
    np.testing.assert_array_equal(relaxed_gamma_window((nx, ny), (0, 0), (nx + 1, ny + 1), rel), g)
    # a window is the slice of the global matrix
    w = relaxed_gamma_window((nx, ny), (5, 3), (9, 8), rel)
    np.testing.assert_array_equal(w, g[5:14, 3:11])


# ------------------------------------------------------------------ two gloo ranks
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, shape):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nz = 2
        for (NX, NY, px, py) in shape:
            d = Decomposition(NX, NY, px, py)
            fields = _local_fields(d, rank, nz, fill_halo=False)
            HaloExchange(d, rank, nz, NF, "cpu").exchange(fields)
            _check(d, rank, fields, nz)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_exchange_two_gloo_ranks():
    import torch.multiprocessing as mp

    shapes = [(18, 9, 2, 1), (9, 18, 1, 2)]
    mp.spawn(_gloo_worker, args=(2, _free_port(), shapes), nprocs=2, join=True)


# ------------------------------------------------------------------ peer-to-peer transport: host logic
@pytest.mark.parametrize("NX,NY,px,py", [(16, 9, 2, 1), (9, 16, 1, 2), (19, 17, 2, 2), (40, 23, 4, 2), (27, 30, 3, 3)])
def test_single_phase_plan_fills_all_halos(NX, NY, px, py):
    """The one-phase plan of the peer-store transport (faces over the owned rows / columns + corner
    blocks from the diagonal neighbours) leaves exactly the halo contents of the two-phase plan."""
    from tasmania_b200.distributed import _OPPOSITE

    nz = 3
    d = Decomposition(NX, NY, px, py)
    fields = [_local_fields(d, r, nz, fill_halo=False) for r in range(d.world)]
    plans = [d.sides_single_phase(r) for r in range(d.world)]
    for r, plan in enumerate(plans):
        names = [s.name for s in plan]
        assert len(names) == len(set(names)) <= 8
        for s in plan:
            peer = [t for t in plans[s.neighbour] if t.name == _OPPOSITE[s.name]]
            assert len(peer) == 1 and peer[0].neighbour == r and peer[0].extent == s.extent
            (si, sj), (di, dj) = peer[0].send_origin, s.extent
            ri, rj = s.recv_origin
            for dst, src in zip(fields[r], fields[s.neighbour]):
                dst.t[ri:ri + di, rj:rj + dj, :nz] = src.t[si:si + di, sj:sj + dj, :nz]
    for r in range(d.world):
        _check(d, r, fields[r], nz)


@pytest.mark.parametrize("phases", [1, 2])
@pytest.mark.parametrize("px,py", [(2, 1), (2, 2), (4, 2), (3, 3)])
def test_p2p_exchange_descriptors_pair_up(monkeypatch, px, py, phases):
    """The tb200_halo_side descriptors of the NVLink transport, built with fake base addresses
    (no GPU): what a rank pushes towards a side lands in the OPPOSITE side's receive buffer and
    counter of exactly that neighbour, slot sizes and extents agree, and the geometry of the slab
    is the one of the message-based exchange."""
    from tasmania_b200 import distributed as D
    from tasmania_b200 import lib

    class FakeLib:
        count = 0

        def tb200_p2p_alloc(self, nbytes, ref):
            FakeLib.count += 1
            ref._obj.value = 0x40000000 * FakeLib.count
            return 0

    monkeypatch.setattr(lib, "_lib", FakeLib())
    nz, nf = 6, 5
    d = Decomposition(24 * px, 20 * py, px, py)
    ex = [D.P2PHaloExchange(d, r, nz, nf, phases=phases) for r in range(d.world)]
    for e in ex:
        e.connect_in_process(ex)

    def plan_of(rank, phase):
        if phases == 2:
            return d.sides(rank, phase)
        return d.sides_single_phase(rank) if phase == 0 else []

    seen = set()
    for e in ex:
        spans = []
        for phase, sides in enumerate(e.plan):
            assert [s.name for s in sides] == [s.name for s in plan_of(e.rank, phase)]
            for m, s in enumerate(sides):
                h, peer, opp = e.sides_c[phase][m], ex[s.neighbour], D._OPPOSITE[s.name]
                assert h.remote_buffer == peer.base + peer.layout[opp]["buffer"]
                assert h.remote_counter == peer.base + peer.layout[opp]["counter"]
                assert h.local_buffer == e.base + e.layout[s.name]["buffer"]
                assert h.local_counter == e.base + e.layout[s.name]["counter"]
                assert h.slot_doubles == nf * nz * s.extent[0] * s.extent[1] == peer.layout[opp]["slot_doubles"]
                assert (tuple(h.send_origin), tuple(h.recv_origin), tuple(h.extent)) == \
                    (s.send_origin, s.recv_origin, s.extent)
                # the slab sent is the window the neighbour receives (global indices)
                peer_side = [t for t in plan_of(s.neighbour, phase) if t.name == opp][0]
                g0, g1 = d.local(e.rank), d.local(s.neighbour)
                assert g0[0] + s.send_origin[0] == g1[0] + peer_side.recv_origin[0]
                assert g0[2] + s.send_origin[1] == g1[2] + peer_side.recv_origin[1]
                assert (h.remote_buffer, h.remote_counter) not in seen  # one writer per buffer
                seen.add((h.remote_buffer, h.remote_counter))
                spans.append((e.layout[s.name]["buffer"], e.layout[s.name]["buffer"] + 16 * h.slot_doubles))
                assert e.layout[s.name]["counter"] % 8 == 0
        spans.sort()
        assert all(a[1] <= b[0] for a, b in zip(spans[:-1], spans[1:]))  # two slots each, no overlap
        assert spans[-1][1] <= min(e.layout[n]["counter"] for n in e.layout)
        assert e.channel_offset[1] + lib.P2P_CHANNEL_BYTES <= e.nbytes
