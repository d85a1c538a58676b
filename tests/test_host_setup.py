# -*- coding: utf-8 -*-
"""CPU checks of the host side: set-up formulas vs the reference fixture, storage wrapper
semantics, registry behaviour, and that the C-ABI library loads and exports every symbol
include/tasmania_b200.h declares (no compute without a GPU)."""
import copy
import os
import re
from datetime import timedelta

import numpy as np
import pytest

import tasmania_b200 as tb
from tasmania_b200 import lib
from tasmania_b200.grid import Grid, Topography, gaussian_profile, isentropic_state_from_brunt_vaisala
from tests import helpers as hp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_setup_matches_reference_fixture():
    fx = hp.load("isen_dry_rk3_5th")
    nx, ny, nz = (int(v) for v in fx["dims"][:3])
    x = np.linspace(-176.0, 176.0, nx)
    y = np.linspace(-176.0, 176.0, ny)
    np.testing.assert_array_equal(gaussian_profile(x, y, 500.0, 50.0, 50.0), fx["topo_steady"])
    grid = Grid((-176.0, 176.0), nx, (-176.0, 176.0), ny, (400.0, 280.0), nz, units_to_m=1e3,
                topography=Topography(fx["topo_steady"], timedelta(seconds=60)))
    np.testing.assert_array_equal(grid.x, fx["x"])
    np.testing.assert_array_equal(grid.z, fx["z"])
    assert grid.dx == float(fx["params"][0]) and grid.dz == float(fx["params"][2])
    st = isentropic_state_from_brunt_vaisala(grid, 22.5, 0.0, 0.015)
    for n in (hp.S, hp.SU, hp.SV, hp.U, hp.V, hp.MTG, hp.P, hp.EXN, hp.H):
        np.testing.assert_array_equal(st[n], fx["init_" + n])


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tasmania_b200.h")).read()
    declared = set(re.findall(r"\b(tb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    l = lib.load()
    for name in declared:
        assert hasattr(l, name), f"{name} declared in the header but not exported"
    assert declared == set(lib.exported_symbols())
    assert l.tb200_version() >= 100


def test_no_cpu_fallback():
    a = np.zeros((4, 4, 2))
    with pytest.raises(tb.B200Error):
        tb.compile_stencil("copy")(src=a, dst=a, origin=(0, 0, 0), domain=(4, 4, 2))
    with pytest.raises(tb.FactoryRegistryError):
        tb.compile_stencil("copy", backend="numpy")
    with pytest.raises(tb.FactoryRegistryError):
        tb.compile_stencil("not_a_stencil")


def test_storage_wrapper_semantics():
    a = tb.zeros((5, 4, 3), device="cpu")
    assert a.shape == (5, 4, 3) and a.dtype == np.float64
    assert a.strides == (8, 16 * 8, 16 * 4 * 8)  # i fastest, rows padded to 16 doubles
    a[1:3, :, 0] = 2.0
    a[:, :, 1] = np.arange(20.0).reshape(5, 4)
    a[...] = a  # self assignment is a no-op
    b = tb.as_storage(np.arange(3.0)[None, None, :], device="cpu")
    a[:2, :2, :] = b  # broadcast
    ref = np.zeros((5, 4, 3))
    ref[1:3, :, 0] = 2.0
    ref[:, :, 1] = np.arange(20.0).reshape(5, 4)
    ref[:2, :2, :] = np.arange(3.0)
    np.testing.assert_array_equal(tb.to_numpy(a), ref)
    v = a[:, :, 1:2]
    assert v.shape == (5, 4, 1) and v.strides == a.strides
    assert a[4, 3, 1] == 19.0
    c = copy.deepcopy(a)
    c[0, 0, 0] = -1.0
    assert a[0, 0, 0] != -1.0
    np.testing.assert_array_equal(np.asarray(c)[1:], ref[1:])
