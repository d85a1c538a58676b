# -*- coding: utf-8 -*-
"""SURVEY.md 8f-1 on the GPU: ``tb200_vertical_advection`` through the compiled b200 stencil and
the host mirror of ``IsentropicVerticalAdvection`` against the reference's own numpy outputs
(tests/golden/isentropic_physics.npz) and, on a larger seeded case, against the oracle.  No
transcendental call in this stencil: every comparison is bit for bit."""
import numpy as np
import pytest

from tests import helpers as hp

pytestmark = pytest.mark.gpu


def _cases():
    for scheme in ("upwind", "centered", "third_order_upwind", "fifth_order_upwind"):
        for z in (0, 1):
            for m in (0, 1):
                for ow in (1, 0):
                    yield scheme, bool(z), bool(m), bool(ow)


def test_stencil_vs_reference_fixture_bitwise():
    import tasmania_b200 as tb
    from tasmania_b200.stencils import FLUX

    fx = hp.load("isentropic_physics")
    nx, ny, nz = (int(v) for v in fx["dims"])
    dz = float(fx["dz"][0])
    dev = {k: tb.as_storage(fx["in_" + k]) for k in ("w", "s", "su", "sv", "qv", "qc", "qr")}
    for scheme, stg, moist, ow in _cases():
        bo = tb.BackendOptions()
        bo.externals = {"get_flux_dry": FLUX[scheme], "get_flux_moist": FLUX[scheme], "moist": moist,
                        "staggering": stg, "flux_extent": FLUX[scheme].extent}
        st = tb.compile_stencil("vertical_advection", backend_options=bo)
        names = ("s", "su", "sv") + (("qv", "qc", "qr") if moist else ())
        outs = {k: tb.as_storage(np.full(fx["in_s"].shape, 7.0) if ow else fx["prev_" + k]) for k in names}
        kw = {"in_w": dev["w"], "dz": dz, "origin": (0, 0, 0), "domain": (nx, ny, nz),
              "exec_info": None, "validate_args": False}
        for k in names:
            kw["in_" + k], kw["out_" + k], kw["ow_out_" + k] = dev[k], outs[k], ow
        st(**kw)
        prefix = f"{scheme}_z{int(stg)}_m{int(moist)}_o{int(ow)}_"
        for k in names:
            np.testing.assert_array_equal(tb.to_numpy(outs[k]), fx[prefix + k], err_msg=prefix + k)


@pytest.mark.parametrize("scheme,moist,stg", [("fifth_order_upwind", True, False),
                                              ("third_order_upwind", False, True),
                                              ("upwind", True, True)])
def test_host_mirror_vs_oracle_bitwise(scheme, moist, stg):
    """The mirror of IsentropicVerticalAdvection.array_call on a 67 x 45 x 60 state (config 3's
    number of levels) with a sub-box origin exercised through the oracle signature."""
    import tasmania_b200 as tb
    from oracle import isentropic_physics as ova
    from tasmania_b200 import isentropic_physics as va
    from tasmania_b200.grid import Grid

    nx, ny, nz = 67, 45, 60
    rng = np.random.default_rng(5)
    shape = (nx + 1, ny + 1, nz + 1)
    grid = Grid((-10.0, 10.0), nx, (-7.0, 7.0), ny, (400.0, 280.0), nz, units_to_m=1e3)
    state = {va.S: rng.uniform(10, 1000, shape), va.SU: rng.uniform(-5e4, 5e4, shape),
             va.SV: rng.uniform(-5e4, 5e4, shape), va.W_ML: rng.uniform(-0.02, 0.02, shape),
             va.W_HL: rng.uniform(-0.02, 0.02, shape), va.MFWV: rng.uniform(0, 5, shape),
             va.MFCW: rng.uniform(0, 5, shape), va.MFPW: rng.uniform(0, 5, shape)}
    names = (va.S, va.SU, va.SV) + ((va.MFWV, va.MFCW, va.MFPW) if moist else ())
    prev = {k: rng.uniform(-1, 1, shape) for k in names}
    ow = {k: (n % 2 == 0) for n, k in enumerate(names)}  # mixed overwrite / accumulate
    comp = va.IsentropicVerticalAdvection(
        grid, scheme, moist=moist, tendency_of_air_potential_temperature_on_interface_levels=stg)
    dstate = {k: tb.as_storage(v) for k, v in state.items()}
    dout = {k: tb.as_storage(prev[k]) for k in names}
    comp.array_call(dstate, dout, {}, ow)
    ref = {k: prev[k].copy() for k in names}
    kw = dict(dz=grid.dz, origin=(0, 0, 0), domain=(nx, ny, nz), ow_out_s=ow[va.S],
              ow_out_su=ow[va.SU], ow_out_sv=ow[va.SV])
    if moist:
        kw.update(in_qv=state[va.MFWV], in_qc=state[va.MFCW], in_qr=state[va.MFPW],
                  out_qv=ref[va.MFWV], out_qc=ref[va.MFCW], out_qr=ref[va.MFPW],
                  ow_out_qv=ow[va.MFWV], ow_out_qc=ow[va.MFCW], ow_out_qr=ow[va.MFPW])
    ova.vertical_advection(scheme, stg, state[va.W_HL] if stg else state[va.W_ML], state[va.S],
                           state[va.SU], state[va.SV], ref[va.S], ref[va.SU], ref[va.SV], **kw)
    for k in names:
        np.testing.assert_array_equal(tb.to_numpy(dout[k]), ref[k], err_msg=k)


def test_uniform_column_has_no_tendency_full_size():
    """Size-independent property at config 5's size (1024 x 1024 x 64): a vertically uniform
    field advected by any velocity has F[k+1] - F[k] = (w[k+1] - w[k]) phi exactly for the
    centered scheme, and zero tendency where w is uniform as well."""
    import torch

    import tasmania_b200 as tb
    from tasmania_b200.stencils import FLUX

    nx, ny, nz = 1024, 1024, 64
    shape = (nx + 1, ny + 1, nz + 1)
    bo = tb.BackendOptions()
    bo.externals = {"get_flux_dry": FLUX["fifth_order_upwind"], "moist": False, "staggering": True}
    st = tb.compile_stencil("vertical_advection", backend_options=bo)
    w = tb.as_storage(np.full(shape, 0.0125))
    s = tb.zeros(shape)
    s.t.copy_(torch.rand(shape[0], shape[1], 1, dtype=torch.float64, device=s.t.device).expand(shape) + 1.0)
    su, sv = tb.as_storage(np.full(shape, -3.0)), tb.as_storage(np.full(shape, 11.0))
    outs = [tb.as_storage(np.full(shape, 5.0)) for _ in range(3)]
    st(in_w=w, in_s=s, in_su=su, in_sv=sv, out_s=outs[0], out_su=outs[1], out_sv=outs[2], dz=2.0,
       ow_out_s=True, ow_out_su=True, ow_out_sv=True, origin=(0, 0, 0), domain=(nx, ny, nz))
    for o in outs:
        assert float(o.t.abs().max()) <= 1e-12  # 60 w phi / 60 - ... cancels up to rounding of w / 60


def test_coriolis_vs_reference_fixture_bitwise():
    """SURVEY.md 8f-3 through the host mirror of IsentropicConservativeCoriolis (nb = 2)."""
    import tasmania_b200 as tb
    from tasmania_b200 import isentropic_physics as va
    from tasmania_b200.grid import Grid

    fx = hp.load("isentropic_physics")
    nx, ny, nz = (int(v) for v in fx["dims"])
    grid = Grid((-10.0, 10.0), nx, (-7.0, 7.0), ny, (400.0, 280.0), nz, units_to_m=1e3)
    comp = va.IsentropicConservativeCoriolis(grid, 2, coriolis_parameter=float(fx["f"][0]))
    state = {va.SU: tb.as_storage(fx["in_su"]), va.SV: tb.as_storage(fx["in_sv"])}
    for owu, owv in ((True, True), (False, True), (False, False)):
        out = {va.SU: tb.as_storage(fx["prev_su"]), va.SV: tb.as_storage(fx["prev_sv"])}
        comp.array_call(state, out, {}, {va.SU: owu, va.SV: owv})
        np.testing.assert_array_equal(tb.to_numpy(out[va.SU]), fx[f"coriolis_o{int(owu)}{int(owv)}_su"])
        np.testing.assert_array_equal(tb.to_numpy(out[va.SV]), fx[f"coriolis_o{int(owu)}{int(owv)}_sv"])


def test_smagorinsky_vs_reference_fixture_bitwise():
    """SURVEY.md 8f-3 through the compiled b200 stencils (generic and isentropic form)."""
    import tasmania_b200 as tb

    fx = hp.load("isentropic_physics")
    nx, ny, nz = (int(v) for v in fx["dims"])
    dx, dy, cs = (float(v) for v in fx["smag"])
    st2 = tb.compile_stencil("smagorinsky")
    sti = tb.compile_stencil("smagorinsky_isentropic")
    d = {k: tb.as_storage(fx["in_" + k]) for k in ("u", "v", "s", "su", "sv")}
    for nb, ow in ((2, True), (3, False)):
        box = dict(origin=(nb, nb, 0), domain=(nx - 2 * nb, ny - 2 * nb, nz))
        a, b = tb.as_storage(fx["prev_su"]), tb.as_storage(fx["prev_sv"])
        st2(in_u=d["u"], in_v=d["v"], out_u_tnd=a, out_v_tnd=b, dx=dx, dy=dy, cs=cs, ow_out_u_tnd=ow,
            ow_out_v_tnd=not ow, **box)
        np.testing.assert_array_equal(tb.to_numpy(a), fx[f"smag2d_nb{nb}_u"])
        np.testing.assert_array_equal(tb.to_numpy(b), fx[f"smag2d_nb{nb}_v"])
        a, b = tb.as_storage(fx["prev_su"]), tb.as_storage(fx["prev_sv"])
        sti(in_s=d["s"], in_su=d["su"], in_sv=d["sv"], out_su_tnd=a, out_sv_tnd=b, dx=dx, dy=dy, cs=cs,
            ow_out_su_tnd=ow, ow_out_sv_tnd=not ow, **box)
        np.testing.assert_array_equal(tb.to_numpy(a), fx[f"smagisen_nb{nb}_su"])
        np.testing.assert_array_equal(tb.to_numpy(b), fx[f"smagisen_nb{nb}_sv"])


@pytest.mark.parametrize("nx,ny,nz,nb", [(131, 77, 9, 2), (64, 40, 3, 3)])
def test_smagorinsky_mirror_vs_oracle_bitwise(nx, ny, nz, nb):
    """Tile seams (several 32 x 8 blocks), ragged edges, several levels: host mirror of
    IsentropicSmagorinsky against the oracle."""
    import tasmania_b200 as tb
    from oracle import isentropic_physics as ova
    from tasmania_b200 import isentropic_physics as va
    from tasmania_b200.grid import Grid

    rng = np.random.default_rng(nx)
    shape = (nx + 1, ny + 1, nz + 1)
    grid = Grid((-100.0, 100.0), nx, (-70.0, 70.0), ny, (400.0, 280.0), nz, units_to_m=1e3)
    s, su, sv = rng.uniform(10, 1000, shape), rng.uniform(-5e4, 5e4, shape), rng.uniform(-5e4, 5e4, shape)
    prev = {va.SU: rng.uniform(-1, 1, shape), va.SV: rng.uniform(-1, 1, shape)}
    comp = va.IsentropicSmagorinsky(grid, nb, smagorinsky_constant=0.21)
    out = {k: tb.as_storage(v) for k, v in prev.items()}
    comp.array_call({va.S: tb.as_storage(s), va.SU: tb.as_storage(su), va.SV: tb.as_storage(sv)}, out, {},
                    {va.SU: False, va.SV: True})
    ra, rb = prev[va.SU].copy(), prev[va.SV].copy()
    ova.smagorinsky(su, sv, ra, rb, in_s=s, dx=grid.dx, dy=grid.dy, cs=0.21, ow_out_u_tnd=False,
                    ow_out_v_tnd=True, origin=(nb, nb, 0), domain=(nx - 2 * nb, ny - 2 * nb, nz))
    np.testing.assert_array_equal(tb.to_numpy(out[va.SU]), ra)
    np.testing.assert_array_equal(tb.to_numpy(out[va.SV]), rb)


def test_implicit_vertical_advection_vs_reference_fixture_bitwise():
    """SURVEY.md 8f-4 through the compiled b200 stencil."""
    import tasmania_b200 as tb

    fx = hp.load("isentropic_physics")
    nx, ny, nz = (int(v) for v in fx["dims"])
    d = {k: tb.as_storage(fx["in_" + k]) for k in ("s", "su", "sv", "qv", "qc", "qr")}
    w = tb.as_storage(fx["in_w_implicit"])
    for z in (0, 1):
        for m in (0, 1):
            bo = tb.BackendOptions()
            bo.externals = {"moist": bool(m), "staggering": bool(z)}
            st = tb.compile_stencil("implicit_vertical_advection", backend_options=bo)
            names = ("s", "su", "sv") + (("qv", "qc", "qr") if m else ())
            outs = {n: tb.as_storage(fx["prev_" + n]) for n in names}
            kw = {"in_w": w, "gamma": float(fx["gamma"][0]), "origin": (0, 0, 0), "domain": (nx, ny, nz)}
            for n in names:
                kw["in_" + n], kw["out_" + n] = d[n], outs[n]
            st(**kw)
            for n in names:
                np.testing.assert_array_equal(tb.to_numpy(outs[n]), fx[f"implicit_z{z}_m{m}_{n}"],
                                              err_msg=f"{z}{m}{n}")


@pytest.mark.parametrize("nz", [60, 100])
def test_implicit_vertical_advection_mirror_vs_oracle_bitwise(nz):
    """Config 3's number of levels (register-sized arrays) and a taller column (the 256-level
    instantiation) through the host mirror, moist, velocity on interface levels."""
    from datetime import timedelta

    import tasmania_b200 as tb
    from oracle import isentropic_physics as ova
    from tasmania_b200 import isentropic_physics as va
    from tasmania_b200.grid import Grid

    nx, ny = 45, 37
    rng = np.random.default_rng(nz)
    shape = (nx + 1, ny + 1, nz + 1)
    grid = Grid((-10.0, 10.0), nx, (-7.0, 7.0), ny, (400.0, 280.0), nz, units_to_m=1e3)
    state = {va.S: rng.uniform(10, 1000, shape), va.SU: rng.uniform(-5e4, 5e4, shape),
             va.SV: rng.uniform(-5e4, 5e4, shape), va.W_HL: rng.uniform(-0.15, 0.15, shape),
             va.MFWV: rng.uniform(0, 5, shape), va.MFCW: rng.uniform(0, 5, shape),
             va.MFPW: rng.uniform(0, 5, shape)}
    names = (va.S, va.SU, va.SV, va.MFWV, va.MFCW, va.MFPW)
    comp = va.IsentropicImplicitVerticalAdvectionDiagnostic(
        grid, moist=True, tendency_of_air_potential_temperature_on_interface_levels=True)
    out = {k: tb.zeros(shape) for k in names}
    dt = timedelta(seconds=12)
    comp.array_call({k: tb.as_storage(v) for k, v in state.items()}, dt, {}, out)
    ref = {k: np.zeros(shape) for k in names}
    ova.implicit_vertical_advection(
        True, state[va.W_HL], state[va.S], state[va.SU], state[va.SV], ref[va.S], ref[va.SU], ref[va.SV],
        gamma=dt.total_seconds() / (4.0 * grid.dz), origin=(0, 0, 0), domain=(nx, ny, nz),
        in_qv=state[va.MFWV], in_qc=state[va.MFCW], in_qr=state[va.MFPW], out_qv=ref[va.MFWV],
        out_qc=ref[va.MFCW], out_qr=ref[va.MFPW])
    for k in names:
        np.testing.assert_array_equal(tb.to_numpy(out[k]), ref[k], err_msg=k)


def test_implicit_vertical_advection_tendency_vs_reference_fixture_bitwise():
    """The prognostic variant through its host mirror (tendencies in component-owned storages)."""
    from datetime import timedelta

    import tasmania_b200 as tb
    from tasmania_b200 import isentropic_physics as va
    from tasmania_b200.grid import Grid

    fx = hp.load("isentropic_physics")
    nx, ny, nz = (int(v) for v in fx["dims"])
    dz = float(fx["dz"][0])
    grid = Grid((-10.0, 10.0), nx, (-7.0, 7.0), ny, (280.0 + dz * nz, 280.0), nz, units_to_m=1e3)
    assert grid.dz == dz
    key = {va.S: "s", va.SU: "su", va.SV: "sv", va.MFWV: "qv", va.MFCW: "qc", va.MFPW: "qr"}
    for z, m in ((0, 1), (1, 0)):
        comp = va.IsentropicImplicitVerticalAdvectionPrognostic(
            grid, moist=bool(m), tendency_of_air_potential_temperature_on_interface_levels=bool(z))
        state = {n: tb.as_storage(fx["in_" + k]) for n, k in key.items()}
        state[va.W_HL if z else va.W_ML] = tb.as_storage(fx["in_w_implicit"])
        tnd, _ = comp.array_call(state, timedelta(seconds=7.5))
        for n, arr in tnd.items():
            got = tb.to_numpy(arr)[:nx, :ny, :nz]
            np.testing.assert_array_equal(got, fx[f"implicit_tnd_z{z}_m{m}_{key[n]}"][:nx, :ny, :nz], err_msg=n)


def test_marching_vertical_advection_variant_bitwise():
    """``TB200_VADV_IMPL=march`` (thread per column and chunk of levels, register windows; measured
    slower than the default at configs[2], kept as a variant): the vertical-advection tests of this
    file and the fused advection step of the moist model give the same bits.  The switch is read
    once per process, hence the subprocess."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider",
           os.path.join(root, "tests", "test_gpu_isentropic_physics.py"), "-k",
           "stencil_vs_reference_fixture or host_mirror_vs_oracle or uniform_column",
           os.path.join(root, "tests", "test_gpu_moist_model.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root,
                         env=dict(os.environ, TB200_VADV_IMPL="march"))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
