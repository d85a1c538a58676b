# -*- coding: utf-8 -*-
"""tb200_ctx: the fused kernels' scratch lives behind an explicit handle held
by the host-side dycore object -- b200 storage layout, zero-filled, the same fields on every request
for a shape, and the dycore that uses them gives the bits of one that uses plain storages."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_context_scratch_layout_and_reuse():
    import torch

    from tasmania_b200 import lib

    ctx = lib.Context()
    a = ctx.scratch((37, 21, 9), 3)
    b = ctx.scratch((37, 21, 9), 2)
    c = ctx.scratch((40, 21, 9), 1)
    assert [x._ptr for x in b] == [x._ptr for x in a[:2]] and c[0]._ptr not in [x._ptr for x in a]
    for f in a:
        cai = f.__cuda_array_interface__
        assert cai["shape"] == (37, 21, 9) and cai["strides"] == (8, 48 * 8, 48 * 21 * 8)
        assert cai["data"][0] % 256 == 0
        t = torch.as_tensor(f, device="cuda")
        assert float(t.abs().sum()) == 0.0
    fld = lib.as_field(a[0])
    assert tuple(fld.shape) == (37, 21, 9) and tuple(fld.stride) == (1, 48, 48 * 21)


def test_dycore_with_context_scratch_equals_plain_storages(monkeypatch):
    from tasmania_b200.distributed import InProcessDecomposedRun

    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("TB200_CTX_SCRATCH", flag)
        run = InProcessDecomposedRun(67, 45, 12, 1, 1, damp_depth=4, topo_seconds=20.0)
        assert hasattr(run.subs[0].dyc, "_ctx") == (flag == "1")
        for _ in range(3):
            run.step()
        res.append({n: run.gather(n) for n in run.subs[0].names})
    for n, v in res[0].items():
        np.testing.assert_array_equal(v, res[1][n], err_msg=n)
