# -*- coding: utf-8 -*-
"""The UNMODIFIED reference classes on backend "b200" against the REAL library on a GPU (VERDICT
round 1, item 5): ten RK3WS steps of the reference's own dry dynamical-core stage
(``IsentropicDynamicalCore.stage_array_call_dry`` with its Domain, Relaxed boundary, prognostic,
damper, velocity and diagnostics objects) through the plugin -- the fused stage behind the
reference's class, and the per-stencil kernels -- equal the same objects on the reference's numpy
backend within north_star's 1e-12.

Needs the reference's sources: staged by ``baseline/stage_reference.sh`` into the git-ignored
``baseline/_ref`` (which travels to the GPU box); skipped when absent."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_root():
    for cand in (os.environ.get("TASMANIA_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "src", "tasmania")):
            return cand
    return None


@pytest.mark.parametrize("mode", ["fused", "per-stencil", "fused-moist", "per-stencil-moist",
                                  "fused-periodic", "per-stencil-periodic", "fused-moist-periodic"])
def test_reference_dycore_on_b200_equals_its_numpy_backend(mode):
    ref = _reference_root()
    if ref is None:
        pytest.skip("reference sources not staged (bash baseline/stage_reference.sh)")
    cmd = [sys.executable, os.path.join(ROOT, "tests", "ref_dycore_steps.py"), "--steps", "10"]
    if mode.startswith("per-stencil"):
        cmd.append("--per-stencil")
    if "moist" in mode:
        cmd.append("--moist")
    if mode.endswith("periodic"):  # the reference's Periodic boundary (tb200_isentropic_stage.periodic when fused)
        cmd.append("--periodic")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, TASMANIA_REFERENCE=ref))
    print(res.stdout[-2000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert f"REF-DYCORE-STEPS-OK {mode} 10" in res.stdout
