# -*- coding: utf-8 -*-
"""The b200 plugin against the UNMODIFIED reference (only where /root/reference is mounted,
i.e. in the build container; skipped on the GPU box)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TASMANIA_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "tasmania")),
                    reason="reference tree not mounted")
def test_plugin_registers_into_the_reference():
    res = subprocess.run([sys.executable, os.path.join(HERE, "plugin_check.py")],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "PLUGIN-OK" in res.stdout


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "tasmania")),
                    reason="reference tree not mounted")
def test_reference_dycore_on_b200_issues_the_mirror_call_sequence():
    """north_star: "IsentropicDynamicalCore ... work unchanged" on the b200 backend."""
    res = subprocess.run([sys.executable, os.path.join(HERE, "ref_dycore_check.py")],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "REF-DYCORE-OK 42 69" in res.stdout  # dry and moist stage sequences
    assert "FUSED-HOOK-OK" in res.stdout        # the fused stage behind the reference's class
    assert "FUSED-MOIST-HOOK-OK" in res.stdout  # ... and the fused moist stage (stage_array_call_moist)
    assert "FUSED-TENDENCIES-HOOK-OK" in res.stdout  # ... and slow tendencies through the fused dry stage


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "tasmania")),
                    reason="reference tree not mounted")
def test_reference_dycore_time_stepped_on_b200_equals_its_numpy_backend():
    """Four RK3WS steps of the reference's own dycore stage, Relaxed boundary, damper and
    diagnostics on backend b200 through plugin.install(fused_stage=True) -- the oracle-backed stub
    standing in for the library -- against the same objects on backend numpy: bit for bit.  (The
    same script runs against the real library in tests/test_gpu_plugin_reference.py.)"""
    res = subprocess.run([sys.executable, os.path.join(HERE, "ref_dycore_steps.py"), "--stub", "--steps", "4"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "REF-DYCORE-STEPS-OK fused 4" in res.stdout


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "tasmania")),
                    reason="reference tree not mounted")
def test_reference_dycore_with_periodic_boundary_on_b200_equals_its_numpy_backend():
    """The same with the reference's Periodic boundary: the plugin's hook runs the fused dry stage
    with ``tb200_isentropic_stage.periodic`` (the wrap of s inside the call, enforce_raw and the
    damping after it, dycore.py:L684-L700), three fused calls per step, the reference's bits."""
    for extra, tag in ((), "fused-periodic"), (("--moist",), "fused-moist-periodic"):
        res = subprocess.run([sys.executable, os.path.join(HERE, "ref_dycore_steps.py"), "--stub", "--periodic",
                              "--steps", "3", "--nx", "23", "--ny", "19", "--nz", "8", *extra],
                             capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stdout + res.stderr
        assert f"REF-DYCORE-STEPS-OK {tag} 3" in res.stdout


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "tasmania")),
                    reason="reference tree not mounted")
def test_reference_physics_components_on_b200_issue_the_mirror_calls():
    """north_star: "the sympl TendencyComponent/DiagnosticComponent classes ... work unchanged":
    the eleven components of the moist benchmark's physics chain and the Burgers stepper."""
    res = subprocess.run([sys.executable, os.path.join(HERE, "ref_components_check.py")],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "REF-COMPONENTS-OK 13" in res.stdout
