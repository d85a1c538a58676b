# -*- coding: utf-8 -*-
"""The fused stage has interchangeable kernels (selected per process by environment
variables, read once): the s-step + scans as kernel S (thread per column) or kernels A + B,
the momentum step as the register-window kernel, the shared-memory-ring kernels (one or two
columns per lane) or the TMA kernel.  Every combination must give bit-identical fields."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import hashlib, sys
sys.path.insert(0, %r)
import numpy as np
from tasmania_b200.distributed import InProcessDecomposedRun
nz = int(sys.argv[2])
run = InProcessDecomposedRun(131, 77, nz, 1, 1, damp_depth=5, topo_seconds=15.0, flux=sys.argv[1])
for _ in range(3):
    run.step()
h = hashlib.sha256()
sub = run.subs[0]
for name in sub.names:
    a = run.gather(name)
    assert np.isfinite(a).all()
    h.update(np.ascontiguousarray(a).tobytes())
print("DIGEST", h.hexdigest())
""" % ROOT


def _digest(flux, nz=12, **env):
    e = dict(os.environ)
    e.update(env)
    res = subprocess.run([sys.executable, "-c", SCRIPT, flux, str(nz)], capture_output=True, text=True,
                         env=e, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    return [l for l in res.stdout.splitlines() if l.startswith("DIGEST")][0]


@pytest.mark.parametrize("flux", ["fifth_order_upwind", "third_order_upwind", "upwind"])
def test_stage_kernel_variants_are_bitwise_identical(flux):
    ref = _digest(flux)
    assert _digest(flux, TB200_S_IMPL="column") == ref
    assert _digest(flux, TB200_STAGE_IMPL="tma") == ref
    assert _digest(flux, TB200_S_IMPL="column", TB200_STAGE_IMPL="tma") == ref
    assert _digest(flux, TB200_MV_IMPL="window") == ref
    assert _digest(flux, TB200_MV_IMPL="ring") == ref
    # the s-step with two columns per lane (the default has one), the velocities read instead of
    # re-diagnosed in the kernels, and the three block shapes of the momentum kernel
    assert _digest(flux, TB200_A_IMPL="two") == ref
    assert _digest(flux, TB200_LAZY_UV="0") == ref
    assert _digest(flux, TB200_A_IMPL="two", TB200_LAZY_UV="0", TB200_LJ="64") == ref
    assert _digest(flux, TB200_LJ="64", TB200_MV_BLOCK="3x1") == _digest(flux, TB200_LJ="64", TB200_MV_BLOCK="6x1") == ref


@pytest.mark.parametrize("nz", [64, 60, 37, 5])
def test_scan_kernel_variants_are_bitwise_identical(nz):
    """Kernels A + B (register-resident columns: full, ragged and short ones) vs the former
    kernel S (thread per column, pressures parked in memory)."""
    ref = _digest("fifth_order_upwind", nz)
    assert _digest("fifth_order_upwind", nz, TB200_S_IMPL="column") == ref


@pytest.mark.parametrize("nz", [12, 60])
def test_small_grid_kernels_are_bitwise_identical(nz):
    """On small grids the stage runs short strips (LJ = 16 or 8 rows per warp instead of 64) and
    the cooperative column-scan kernel (32 columns x 8 threads per block); the default of this
    131 x 77 case.  Same bits as the large-grid kernels."""
    ref = _digest("fifth_order_upwind", nz)
    assert _digest("fifth_order_upwind", nz, TB200_LJ="64", TB200_B_IMPL="column") == ref
    assert _digest("fifth_order_upwind", nz, TB200_LJ="16", TB200_B_IMPL="coop") == ref
    # the diagnostics refresh: cooperative kernel (default at this size) vs one thread per column
    assert _digest("fifth_order_upwind", nz, TB200_DIAG_IMPL="column") == ref
    assert _digest("fifth_order_upwind", nz, TB200_DIAG_IMPL="coop") == ref
