# -*- coding: utf-8 -*-
"""Domain decomposition on the GPU: the decomposed run (sub-domains in process, exchanged
through the halo pack / unpack kernels) must equal the single-device run BITWISE
(SURVEY.md section 8e), and the single-device sub-domain driver must equal the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fields():
    from tasmania_b200.isentropic import MTG, S, SU, SV, U, V

    return (S, SU, SV, U, V, MTG)


@pytest.mark.parametrize("px,py", [(2, 1), (1, 2), (2, 2), (3, 2)])
def test_decomposed_equals_single_device_bitwise(px, py):
    from tasmania_b200.distributed import InProcessDecomposedRun

    NX, NY, nz = 41, 37, 9
    kw = dict(damp_depth=4, topo_seconds=20.0)
    single = InProcessDecomposedRun(NX, NY, nz, 1, 1, **kw)
    multi = InProcessDecomposedRun(NX, NY, nz, px, py, **kw)
    for _ in range(4):
        single.step()
        multi.step()
    for name in _fields():
        a, b = single.gather(name), multi.gather(name)
        assert np.isfinite(a).all()
        np.testing.assert_array_equal(a, b, err_msg=name)
    # the flow has actually developed (the mountain has grown)
    from tasmania_b200.isentropic import SV

    assert np.abs(single.gather(SV)).max() > 0.0


@pytest.mark.parametrize("phases", ["1", "2"])
@pytest.mark.parametrize("px,py", [(2, 1), (2, 2), (3, 2)])
def test_decomposed_p2p_transport_equals_single_device_bitwise(monkeypatch, px, py, phases):
    """The NVLink peer-store transport (tb200_halo_push / tb200_halo_pull) with all sub-domains in
    one process: every push of a phase is enqueued before the first pull.  One-phase plan (faces +
    corner blocks, the default) and the two-phase plan of the message-based exchange."""
    from tasmania_b200.distributed import InProcessDecomposedRun

    monkeypatch.setenv("TB200_HALO_PHASES", phases)

    NX, NY, nz = 41, 37, 9
    kw = dict(damp_depth=4, topo_seconds=20.0)
    single = InProcessDecomposedRun(NX, NY, nz, 1, 1, **kw)
    multi = InProcessDecomposedRun(NX, NY, nz, px, py, transport="p2p", **kw)
    assert multi.transport == "p2p"
    for _ in range(5):  # both slots of every receive buffer are reused
        single.step()
        multi.step()
    for s in multi.subs:
        s.halo.check()
    for name in _fields():
        np.testing.assert_array_equal(single.gather(name), multi.gather(name), err_msg=name)


def test_p2p_push_pull_roundtrip_and_slots():
    """Two 'ranks' side by side in one process: what rank 0 pushes east is what rank 1 pulls from
    the west, exchange after exchange (slot q & 1), for 3 and for 5 fields."""
    import tasmania_b200 as tb
    from tasmania_b200.distributed import Decomposition, P2PHaloExchange, exchange_in_process_p2p

    nz = 5
    d = Decomposition(40, 21, 2, 1)
    ex = [P2PHaloExchange(d, r, nz, 5) for r in range(2)]
    for e in ex:
        e.connect_in_process(ex)
    rng = np.random.default_rng(11)
    for it, nf in enumerate((5, 3, 5, 3, 3)):
        src, fields = [], []
        for r in range(2):
            nxl, nyl = d.local_shape(r)
            arrs = [rng.standard_normal((nxl + 1, nyl + 1, nz + 1)) for _ in range(nf)]
            src.append(arrs)
            fields.append([tb.as_storage(a) for a in arrs])
        exchange_in_process_p2p(ex, fields)
        got0 = [tb.to_numpy(f) for f in fields[0]]
        got1 = [tb.to_numpy(f) for f in fields[1]]
        h = d.halo
        n0 = d.local_shape(0)[0]
        for n in range(nf):
            # rank 1's west halo = rank 0's last owned columns, and vice versa
            np.testing.assert_array_equal(got1[n][0:h, :21, :nz], src[0][n][n0 - 2 * h:n0 - h, :21, :nz])
            np.testing.assert_array_equal(got0[n][n0 - h:n0, :21, :nz], src[1][n][h:2 * h, :21, :nz])
    for e in ex:
        e.check()
        e.close()


def test_halo_pack_unpack_roundtrip():
    import torch

    import tasmania_b200 as tb
    from tasmania_b200.distributed import _pack, _unpack

    rng = np.random.default_rng(5)
    shape = (23, 19, 7)
    src = [rng.standard_normal(shape) for _ in range(5)]
    fields = [tb.as_storage(a) for a in src]
    origin, extent, nz = (3, 2), (4, 15), 6
    buf = torch.empty(5 * nz * extent[0] * extent[1], dtype=torch.float64, device="cuda")
    _pack(fields, buf, origin, extent, nz)
    got = buf.cpu().numpy().reshape(5, nz, extent[1], extent[0])
    for n in range(5):
        want = src[n][3:7, 2:17, :nz].transpose(2, 1, 0)
        np.testing.assert_array_equal(got[n], want)
    dst = [tb.zeros(shape) for _ in range(5)]
    _unpack(dst, buf, (10, 1), extent, nz)
    for n in range(5):
        out = tb.to_numpy(dst[n])
        np.testing.assert_array_equal(out[10:14, 1:16, :nz], src[n][3:7, 2:17, :nz])
        out[10:14, 1:16, :nz] = 0.0
        assert not out.any()


def test_two_rank_nccl_run_equals_single_device_bitwise():
    """Real multi-process run (torchrun, NCCL) when the box has >= 2 GPUs: tests/mgpu_check.py
    compares every rank's owned block with the single-domain run, with and without
    communication / computation overlap."""
    import os
    import subprocess
    import sys

    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for port, extra in ((29531, ["--transport", "p2p"]), (29532, ["--transport", "p2p", "--overlap"]),
                        (29535, ["--transport", "p2p", "--phases", "2"]),
                        (29533, ["--transport", "nccl"]), (29534, ["--transport", "nccl", "--overlap"])):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(port),
               os.path.join(root, "tests", "mgpu_check.py"), *extra]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
        assert res.returncode == 0 and "MGPU-OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
