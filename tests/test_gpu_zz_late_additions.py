# -*- coding: utf-8 -*-
"""Parity of the one-dimensional diffusers and smoothers (``..._1dx`` / ``..._1dy``;
``tb200_diffusion_1d`` / ``tb200_smoothing_1d``, csrc/horizontal.cu) and of the global ``thomas``
stencil (``tb200_thomas``, csrc/vertical.cu) against (a) the fixture the
reference's own classes wrote (tests/golden/stencils_1d.npz) and (b) the oracle on ragged sizes.
No libm calls: BIT-EXACT.

The file sorts last on purpose: these kernels were added at the end of round 1, after the
round's GPU minutes were spent (built and checked with cuobjdump here, host path checked
numerically through the oracle-backed ABI stub), so their first run on a B200 is the driver's.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import tasmania_b200 as tb  # noqa: E402
from oracle import dwarfs as od  # noqa: E402
from tasmania_b200.dwarfs import HorizontalDiffusion, HorizontalSmoothing  # noqa: E402
from tests import helpers as hp  # noqa: E402

DIFFUSERS = ((2, "second_order"), (4, "fourth_order"))
SMOOTHERS = ((1, "first_order"), (2, "second_order"), (3, "third_order"))


def dev(a):
    return tb.as_storage(np.asarray(a))


def eq(a, b):
    np.testing.assert_array_equal(tb.to_numpy(a), b)


@pytest.mark.parametrize("ax", ("x", "y"))
@pytest.mark.parametrize("tag", ("row", "grid"))
def test_one_dimensional_dwarfs_golden(ax, tag):
    fx = hp.load("stencils_1d")
    dx, dy = fx["scalars"]
    phi = fx[f"{ax}_{tag}_phi"]
    shape = phi.shape
    for order, name in DIFFUSERS:
        diff = HorizontalDiffusion.factory(f"{name}_1d{ax}", shape, dx, dy, 0.5, 1.0, 3)
        tnd = tb.zeros(shape)
        diff(dev(phi), tnd, overwrite_output=True)
        eq(tnd, fx[f"k8_{order}_{ax}_{tag}_tnd"])
        acc = dev(fx[f"{ax}_{tag}_base"])
        diff(dev(phi), acc, overwrite_output=False)
        eq(acc, fx[f"k8_{order}_{ax}_{tag}_acc"])
    for order, name in SMOOTHERS:
        out = tb.zeros(shape)
        HorizontalSmoothing.factory(f"{name}_1d{ax}", shape, 0.03, 0.24, 3)(dev(phi), out)
        eq(out, fx[f"k9_{order}_{ax}_{tag}_out"])


@pytest.mark.parametrize("shape", ((7, 1, 1), (1, 7, 1), (131, 3, 4), (5, 203, 2), (1025, 1, 64)))
def test_one_dimensional_dwarfs_oracle_ragged(shape):
    rng = np.random.default_rng(11 + shape[0] + shape[1])
    phi = rng.standard_normal(shape)
    depth = min(2, shape[2])
    for axis, ax in enumerate("xy"):
        if shape[axis] < 7:
            continue
        for order, name in DIFFUSERS:
            nb = order // 2
            g = np.zeros(shape)
            g[...] = od.vertical_profile(0.5, 1.0, depth, shape[2])[None, None, :]
            origin = tuple(nb if a == axis else 0 for a in range(3))
            domain = tuple(n - 2 * nb if a == axis else n for a, n in enumerate(shape))
            exp = np.zeros(shape)
            od.diffusion_1d(order, axis, phi, g, exp, (0.7, 1.3)[axis], True, origin, domain)
            out = tb.zeros(shape)
            HorizontalDiffusion.factory(f"{name}_1d{ax}", shape, 0.7, 1.3, 0.5, 1.0, depth)(dev(phi), out)
            eq(out, exp)
        for order, name in SMOOTHERS:
            g = np.zeros(shape)
            g[...] = od.vertical_profile(0.03, 0.24, depth, shape[2])[None, None, :]
            exp = np.zeros(shape)
            od.horizontal_smoothing_1d(order, axis, phi, g, exp)
            out = tb.zeros(shape)
            HorizontalSmoothing.factory(f"{name}_1d{ax}", shape, 0.03, 0.24, depth)(dev(phi), out)
            eq(out, exp)


def test_one_dimensional_entry_points_reject_bad_arguments():
    from tasmania_b200 import lib

    shape = (9, 1, 2)
    phi, out = dev(np.ones(shape)), tb.zeros(shape)
    diff = HorizontalDiffusion.factory("fourth_order_1dx", shape, 1.0, 1.0, 0.5, 1.0, 0)
    with pytest.raises(lib.B200Error):  # halo 2 along x does not fit around origin 1
        diff._stencil(in_phi=phi, in_gamma=diff._gamma, out_phi=out, dx=1.0, dy=1.0,
                      ow_out_phi=True, origin=(1, 0, 0), domain=(7, 1, 2))
    with pytest.raises(lib.B200Error):  # in place
        diff._stencil(in_phi=phi, in_gamma=diff._gamma, out_phi=phi, dx=1.0, dy=1.0,
                      ow_out_phi=True, origin=(2, 0, 0), domain=(5, 1, 2))


def test_thomas_golden_and_views():
    from tasmania_b200.framework import BackendOptions

    fx = hp.load("stencils_1d")
    box = [int(v) for v in fx["thomas_box"]]
    thomas = tb.compile_stencil("thomas", backend_options=BackendOptions())
    a, b, c, d = (dev(fx["thomas_" + n]) for n in "abcd")
    x = tb.zeros(fx["thomas_a"].shape)
    thomas(a=a, b=b, c=c, d=d, out=x, origin=tuple(box[:3]), domain=tuple(box[3:]))
    eq(x, fx["thomas_x"])
    sl = (slice(1, 6), slice(0, 5), slice(1, 8))
    dd = dev(fx["thomas_d"])  # views of the storages, solved in place of d
    thomas(a=a[sl], b=b[sl], c=c[sl], d=dd[sl], out=dd[sl], origin=(0, 0, 0), domain=(5, 5, 7))
    np.testing.assert_array_equal(tb.to_numpy(dd)[sl], fx["thomas_x"][sl])


@pytest.mark.parametrize("shape", ((3, 2, 1), (37, 21, 60), (130, 5, 100)))
def test_thomas_oracle_ragged(shape):
    from oracle import isentropic_physics as op
    from tasmania_b200.framework import BackendOptions

    rng = np.random.default_rng(shape[2])
    a, c, d = (rng.uniform(-1, 1, size=shape) for _ in range(3))
    b = rng.uniform(2.5, 4, size=shape) * rng.choice([-1.0, 1.0], size=shape)
    exp = np.zeros(shape)
    op.thomas(a, b, c, d, exp, (0, 0, 0), shape)
    x = tb.zeros(shape)
    tb.compile_stencil("thomas", backend_options=BackendOptions())(
        a=dev(a), b=dev(b), c=dev(c), d=dev(d), out=x, origin=(0, 0, 0), domain=shape)
    eq(x, exp)
    # the solution solves the system (size-independent property, rtol from the conditioning)
    x = tb.to_numpy(x)
    res = b * x
    res[:, :, 1:] += a[:, :, 1:] * x[:, :, :-1]
    res[:, :, :-1] += c[:, :, :-1] * x[:, :, 1:]
    np.testing.assert_allclose(res, d, rtol=0, atol=1e-12)


def test_hyperdiffusion_golden_and_oracle():
    from tasmania_b200.framework import BackendOptions

    fx = hp.load("stencils_1d")
    box = [int(v) for v in fx["hyper_box"]]
    hyper = tb.compile_stencil("hyperdiffusion", backend_options=BackendOptions())
    out = tb.zeros(fx["hyper_phi"].shape)
    hyper(in_phi=dev(fx["hyper_phi"]), out_phi=out, alpha=float(fx["hyper_alpha"]),
          origin=tuple(box[:3]), domain=tuple(box[3:]))
    eq(out, fx["hyper_out"])
    shape = (141, 67, 5)
    phi = np.random.default_rng(3).standard_normal(shape)
    exp = np.zeros(shape)
    od.hyperdiffusion(phi, exp, 0.01, (3, 3, 0), (135, 61, 5))
    out = tb.zeros(shape)
    hyper(in_phi=dev(phi), out_phi=out, alpha=0.01, origin=(3, 3, 0), domain=(135, 61, 5))
    eq(out, exp)


@pytest.mark.parametrize("dims", ((1, 1, 1), (2, 3, 1), (33, 9, 2), (40, 7, 64), (37, 10, 70)))
def test_k3_column_scans_from_one_level_to_deep_columns(dims):
    """The stand-alone K3 stencils at depths the golden grid (6 levels) does not reach: one
    level, and columns beyond the 64 levels of the register-resident scan (the generic kernel).
    Pressure is bit-exact (no libm call); the pow-based outputs within 1e-13 like
    tests/test_gpu_stencils.py."""
    from oracle import isentropic as oi
    from tasmania_b200.framework import BackendOptions

    def stencil(name):
        return tb.compile_stencil(name, backend_options=BackendOptions(externals=dict(oi.CONSTANTS)))

    nx, ny, nz = dims
    shape = (nx + 1, ny + 1, nz + 1)
    rng = np.random.default_rng(nx * 100 + nz)
    theta1d = np.linspace(400.0, 280.0, nz + 1)
    theta = np.zeros(shape)
    theta[:nx, :ny, :] = theta1d[None, None, :]
    hs = np.zeros(shape)
    hs[:nx, :ny, nz] = rng.uniform(0, 800, size=(nx, ny))
    s = rng.uniform(5, 60, size=shape)
    dz, pt = 120.0 / nz, 11868.9
    box = dict(origin=(0, 0, 0), domain=(nx, ny, nz + 1))
    want = [np.zeros(shape) for _ in range(4)]
    oi.diagnostic_variables(theta, hs, s, *want, dz=dz, pt=pt, **box)
    got = [tb.zeros(shape) for _ in range(4)]
    stencil("diagnostic_variables")(in_theta=dev(theta), in_hs=dev(hs), in_s=dev(s), inout_p=got[0],
                                    out_exn=got[1], inout_mtg=got[2], inout_h=got[3], dz=dz, pt=pt, **box)
    eq(got[0], want[0])
    for a, b in zip(got[1:], want[1:]):
        assert hp.relerr(tb.to_numpy(a), b) <= 1e-13
    mtg = tb.zeros(shape)
    stencil("montgomery")(in_hs=dev(hs), in_s=dev(s), inout_mtg=mtg, dz=dz, pt=pt,
                          theta_s=float(theta1d[-1]), **box)
    assert hp.relerr(tb.to_numpy(mtg), want[2]) <= 1e-13
    h, want_h = tb.zeros(shape), np.zeros(shape)  # the stand-alone height scan has its own formula
    oi.height(theta, hs, s, want_h, dz=dz, pt=pt, **box)
    stencil("height")(in_theta=dev(theta), in_hs=dev(hs), in_s=dev(s), inout_h=h, dz=dz, pt=pt, **box)
    assert hp.relerr(tb.to_numpy(h), want_h) <= 1e-13


def test_one_dimensional_boundary_mirrors_golden():
    """Relaxed1DX / 1DY (irelax kernel + slab copies on device storages) and Periodic1DX / 1DY
    against the fixture the reference's own classes wrote."""
    hp.check_one_dimensional_boundaries(hp.load("stencils_1d"), tb)


MARCH_SCRIPT = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
import tasmania_b200 as tb
from oracle import dwarfs as od
from tasmania_b200.dwarfs import HorizontalDiffusion
for shape in ((9, 9, 1), (66, 21, 4), (131, 75, 3), (300, 200, 70)):
    rng = np.random.default_rng(7 + shape[0])
    phi = rng.standard_normal(shape)
    base = rng.standard_normal(shape)
    depth = min(2, shape[2])
    for order, name in ((2, "second_order"), (4, "fourth_order")):
        nb = order // 2
        g = np.zeros(shape)
        g[...] = od.vertical_profile(0.5, 1.0, depth, shape[2])[None, None, :]
        box = ((nb, nb, 0), (shape[0] - 2 * nb, shape[1] - 2 * nb, shape[2]))
        hd = HorizontalDiffusion.factory(name, shape, 0.7, 1.3, 0.5, 1.0, depth)
        for overwrite in (True, False):
            exp = base.copy()
            od.diffusion(order, phi, g, exp, 0.7, 1.3, overwrite, *box)
            out = tb.as_storage(base)
            hd(tb.as_storage(phi), out, overwrite_output=overwrite)
            np.testing.assert_array_equal(tb.to_numpy(out), exp)
print("MARCH-OK")
""" % __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))


def test_marching_diffusion_variant_equals_oracle():
    """TB200_DIFF_IMPL=march (csrc/horizontal.cu:march_kernel, experimental): same bits as the
    oracle -- and hence as the default tiled kernel -- on ragged sizes, both output modes."""
    import os
    import subprocess
    import sys

    env = dict(os.environ, TB200_DIFF_IMPL="march")
    res = subprocess.run([sys.executable, "-c", MARCH_SCRIPT], capture_output=True, text=True,
                         env=env, timeout=300)
    assert res.returncode == 0 and "MARCH-OK" in res.stdout, res.stdout + res.stderr
