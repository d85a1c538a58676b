# -*- coding: utf-8 -*-
"""Run by tests/test_plugin_reference.py in a subprocess (the reference loader patches
sys.modules).  Loads the UNMODIFIED reference from /root/reference through
tests/golden/refload.py, installs the b200 plugin into its registries and drives the
reference's own classes with backend="b200" as far as the CPU container allows: everything
up to (not including) the kernel launches, with storages on the host."""
import inspect
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import refload  # noqa: E402

refload.install_framework()

from tasmania_b200 import plugin, storage  # noqa: E402
from tasmania_b200 import stencils as st  # noqa: E402

storage.DEFAULT_DEVICE_OVERRIDE = "cpu"  # no GPU here: exercise registration + storages only
report = plugin.install()

from tasmania.framework import allocators as ta  # noqa: E402
from tasmania.framework import stencil as ts  # noqa: E402
from tasmania.framework.generic_functions import to_numpy  # noqa: E402
from tasmania.framework.options import BackendOptions, StorageOptions  # noqa: E402
from tasmania.utils.exceptions import FactoryRegistryError  # noqa: E402

# ---- allocators and conversions through the reference's own entry points
z = ta.zeros("b200", shape=(5, 4, 3), storage_options=StorageOptions())
assert isinstance(z, storage.B200Array) and z.shape == (5, 4, 3) and z.dtype == np.float64
o = ta.ones("b200", shape=(2, 2, 2))
assert to_numpy(o).sum() == 8.0 and isinstance(to_numpy(o), np.ndarray)
a = ta.as_storage("b200", data=np.arange(24.0).reshape(2, 3, 4))
assert isinstance(a, storage.B200Array) and np.array_equal(to_numpy(a), np.arange(24.0).reshape(2, 3, 4))

# ---- global stencils resolve to the b200 definitions, compile through the b200 compiler
for name in plugin.GLOBAL_STENCILS:
    assert ts.StencilDefinition("b200", name) is st.framework_definition(name), name
compiled = ts.StencilCompiler("irelax", "b200", backend_options=BackendOptions())
assert callable(compiled) and compiled.externals == {}
try:
    ts.StencilDefinition("b200", "no_such_stencil")  # must fail like any unknown name
    raise SystemExit("an unknown stencil resolved for b200")
except FactoryRegistryError:
    pass

# ---- class-scoped definitions on the reference's classes
must = {"IsentropicDiagnostics:montgomery", "IsentropicDiagnostics:diagnostic_variables",
        "HorizontalVelocity:velocity_x", "WaterConstituent:density", "FourthOrder:diffusion",
        "SecondOrder:smoothing", "SecondOrder1DX:diffusion", "FourthOrder1DY:diffusion",
        "FirstOrder1DX:smoothing", "ThirdOrder1DY:smoothing", "Rayleigh:damping", "BurgersStepper:forward_euler",
        "FifthOrderUpwind:flux_dry", "ThirdOrder:advection"}
missing = must - set(report["class_scoped"])
assert not missing, (missing, report["skipped"])

from tasmania.dwarfs.horizontal_diffusion import HorizontalDiffusion  # noqa: E402
from tasmania.dwarfs.subclasses.horizontal_diffusers import fourth_order, second_order  # noqa: E402,F401

hd = HorizontalDiffusion.factory("fourth_order", (12, 11, 6), 1.0, 1.0, 0.5, 1.0, 3, backend="b200",
                                 backend_options=BackendOptions(), storage_options=StorageOptions())
assert isinstance(hd._gamma, storage.B200Array)           # allocated by the b200 allocator
g = to_numpy(hd._gamma)
assert g.shape == (12, 11, 6) and g[0, 0, 0] > g[0, 0, 5] == 0.5   # broadcast assignment worked
assert hd.get_stencil_definition("diffusion").__name__ == "diffusion_fourth_order_b200"
assert callable(hd._stencil)

from tasmania.isentropic.dynamics.subclasses.minimal_horizontal_fluxes.fifth_order_upwind import (  # noqa: E402
    FifthOrderUpwind)

hf = FifthOrderUpwind(backend="b200")
d = hf.get_subroutine_definition("flux_dry")
assert d.tb200_scheme is st.FLUX["fifth_order_upwind"] and hf.extent == 3
# what RK3WSSI._stencils_initialize does (rk3ws_si.py:L249-L264) -> the b200 K1 compiles
bo = BackendOptions()
bo.externals = {"extent": hf.extent, "flux_dry": d, "flux_moist": hf.get_subroutine_definition("flux_moist"),
                "moist": False, "s_tnd_on": False}
k1 = ts.StencilCompiler("step_forward_euler", "b200", backend_options=bo)
assert st._flux_code(k1.externals) == 3
bo.externals = {}  # the compiler snapshots: later overwrites must not leak in
assert st._flux_code(k1.externals) == 3

# ---- keyword names: every argument of the reference's numpy definition is accepted by ours
import importlib  # noqa: E402

checked = 0
for modname, clsname, stencil, fn in plugin._class_scoped_stencils():
    try:
        cls = getattr(importlib.import_module(modname), clsname)
    except Exception:
        continue
    ref_def = None
    for _, h in inspect.getmembers(cls, predicate=inspect.isfunction):
        tag = getattr(h, "__tasmania__", None)
        if tag and tag.get("function") == "stencil_definition":
            backends, stencils = tag.get("backend"), tag.get("stencil")
            backends = [backends] if isinstance(backends, str) else backends
            stencils = [stencils] if isinstance(stencils, str) else stencils
            if "numpy" in backends and stencil in stencils and not getattr(h, "__isabstractmethod__", False):
                ref_def = h
    assert ref_def is not None, (clsname, stencil)
    ours = set(inspect.signature(fn).parameters) - {"externals"}
    theirs = set(inspect.signature(ref_def).parameters) - {"self"}  # some are instance methods
    assert theirs <= ours, (clsname, stencil, sorted(theirs - ours))
    checked += 1
assert checked >= 14, checked
# ---- the reference's own objects issue the b200 ABI calls (recording stub: every call is
# type-checked against the C signature; fma / copy are carried out on the host buffers)
sys.path.insert(0, ROOT)
from datetime import datetime, timedelta  # noqa: E402
import types  # noqa: E402

from tests.abi_stub import stubbed_library  # noqa: E402

with stubbed_library() as stub:
    phi = ta.as_storage("b200", data=np.random.default_rng(0).standard_normal((12, 11, 6)))
    tnd = ta.zeros("b200", shape=(12, 11, 6))
    hd(phi, tnd)  # tasmania.HorizontalDiffusion.__call__ -> compiled b200 stencil -> C ABI
    assert stub.calls == ["tb200_diffusion"], stub.calls

    # DataArrayDictOperator.fma of the unmodified reference on b200 storages
    from tasmania.utils.xarrayx import DataArrayDictOperator  # noqa: E402

    def da(arr, units="m"):
        return refload.DataArray(arr, None, ("x", "y", "z"), None, {"units": units})

    y0 = np.random.default_rng(1).standard_normal((5, 4, 3))
    op = DataArrayDictOperator(backend="b200", backend_options=BackendOptions(),
                               storage_options=StorageOptions())
    out = {"y": da(ta.zeros("b200", shape=y0.shape))}
    op.fma({"y": da(ta.as_storage("b200", data=y0))}, {"y": da(ta.as_storage("b200", data=2.0 * y0))},
           0.25, out=out)
    # (the plugin batches the per-field fma calls of the operator into one launch)
    assert report["batched"] == ["DataArrayDictOperator.fma", "Relaxed.enforce_raw"]
    assert stub.calls[-1] == "tb200_fma_fields"
    assert np.array_equal(to_numpy(out["y"].data), y0 + 0.25 * (2.0 * y0))
    two = {n: da(ta.as_storage("b200", data=y0 * k)) for k, n in enumerate(("a", "b"), 1)}
    inc = {n: da(ta.as_storage("b200", data=y0 + k)) for k, n in enumerate(("a", "b"), 1)}
    res = {n: da(ta.zeros("b200", shape=y0.shape)) for n in ("a", "b")}
    n0 = stub.count("tb200_fma_fields")
    op.fma(two, inc, -0.5, out=res)
    assert stub.count("tb200_fma_fields") == n0 + 1        # two fields, one launch
    for k, n in enumerate(("a", "b"), 1):
        assert np.array_equal(to_numpy(res[n].data), y0 * k + -0.5 * (y0 + k))
    # other backends keep the reference's own method
    nop = DataArrayDictOperator(backend="numpy")
    nres = nop.fma({"y": da(y0.copy())}, {"y": da(2.0 * y0)}, 0.25, out={"y": da(np.zeros_like(y0))})
    assert np.array_equal(nres["y"].data, y0 + 0.25 * (2.0 * y0))

    # the reference's RK3WS tendency stepper logic (its own _call, unbound) on b200 storages
    from tasmania.framework.subclasses.tendency_steppers.rk3ws import RK3WS  # noqa: E402

    class Increment:
        def get_increment(self, st, timestep, out_increment=None, out_diagnostics=None):
            inc = ta.as_storage("b200", data=-0.7 * to_numpy(st["y"].data))
            return {"y": da(inc), "time": st["time"]}, {}

    fake = types.SimpleNamespace(_stepper_operator=Increment(), _dict_op=op, _enforce_hb=False,
                                 _increment=None, _diagnostics=None,
                                 output_properties={"y": {"units": "m", "dims": ("x", "y", "z")}})
    state = {"y": da(ta.as_storage("b200", data=y0)), "time": datetime(2000, 1, 1)}
    _, out_state = RK3WS._call(fake, state, timedelta(seconds=0.3), {},
                               {"y": da(ta.zeros("b200", shape=y0.shape))})
    z = 0.7 * 0.3
    want = y0 + 0.3 * (-0.7 * (y0 + 0.5 * 0.3 * (-0.7 * (y0 + (1.0 / 3.0 * 0.3) * (-0.7 * y0)))))
    assert np.array_equal(to_numpy(out_state["y"].data), want), np.abs(to_numpy(out_state["y"].data) - want).max()
    assert abs(float(want.ravel()[0] / y0.ravel()[0]) - (1 - z + z * z / 2 - z**3 / 6)) < 1e-15
    assert stub.count("tb200_elementwise") == 0 and stub.count("tb200_fma_fields") >= 3 + 2

    # Relaxed.enforce_raw of the unmodified reference on b200 storages: one frame launch for all
    # fields (carried out on the host by the stub), equal to the reference's numpy backend
    import generate_golden as gg  # noqa: E402

    names = ("air_isentropic_density", "x_velocity_at_u_locations", "y_velocity_at_v_locations",
             "air_pressure_on_interface_levels")
    nx, ny, nz = 19, 17, 5
    rng = np.random.default_rng(3)
    fields = {n: rng.standard_normal((nx + 1, ny + 1, nz + 1)) for n in names}
    refs = {n: rng.standard_normal((nx + 1, ny + 1, nz + 1)) for n in names}
    results = {}
    for backend in ("numpy", "b200"):
        dom = gg._make_domain(nx, ny, nz, "relaxed", 3, {"nr": 6}, topo=False)
        if backend == "b200":  # _make_domain builds numpy objects: rebuild the boundary on b200
            hbm = refload.load("tasmania.domain.horizontal_boundary")
            hb_ = hbm.HorizontalBoundary.factory("relaxed", dom.physical_grid, 3, backend="b200",
                                                 storage_options=StorageOptions(), nr=6)
        else:
            hb_ = dom.horizontal_boundary
        conv = (lambda a: ta.as_storage("b200", data=a)) if backend == "b200" else (lambda a: a.copy())
        hb_.reference_state = {n: refload.DataArray(conv(v), attrs={"units": "1"}) for n, v in refs.items()}
        state = {n: conv(v) for n, v in fields.items()}
        n0 = stub.count("tb200_relax_frame")
        hb_.enforce_raw(state, {n: {"units": "1"} for n in names[:3]})
        if backend == "b200":
            assert stub.count("tb200_relax_frame") == n0 + 1 and stub.count("tb200_relax") == 0
            assert hb_._b200_free_box == (6, nx - 6, 6, ny - 6)
        results[backend] = {n: np.array(to_numpy(v)) for n, v in state.items()}
    for n in names:
        assert np.array_equal(results["b200"][n], results["numpy"][n]), n
    assert np.array_equal(results["b200"][names[3]], fields[names[3]])       # not selected: untouched
    assert not np.array_equal(results["b200"][names[0]], fields[names[0]])

# ---- the reference's one-dimensional diffusers / smoothers on backend b200 (their own __call__:
# the stencil + the rim copies), carried out by the oracle behind the C ABI, against the same
# classes on the numpy backend
from tests.abi_oracle import OracleStub  # noqa: E402
from tasmania.dwarfs.horizontal_smoothing import HorizontalSmoothing  # noqa: E402
from tasmania.dwarfs.subclasses.horizontal_smoothers import first_order as _f1, third_order as _f3  # noqa: E402,F401
from tasmania.dwarfs.subclasses.horizontal_smoothers import second_order as _f2  # noqa: E402,F401

with stubbed_library(OracleStub) as stub:
    rng = np.random.default_rng(5)
    for ax, shape in (("x", (15, 1, 4)), ("y", (1, 16, 4)), ("x", (13, 5, 3)), ("y", (6, 14, 3))):
        phi = rng.standard_normal(shape)
        for name in ("second_order", "fourth_order"):
            res = {}
            for backend in ("numpy", "b200"):
                obj = HorizontalDiffusion.factory(
                    f"{name}_1d{ax}", shape, 1100.0, 900.0, 0.5, 1.0, 2, backend=backend,
                    backend_options=BackendOptions(), storage_options=StorageOptions())
                tnd = ta.zeros(backend, shape=shape)
                obj(ta.as_storage(backend, data=phi), tnd)
                res[backend] = np.array(to_numpy(tnd))
            assert np.array_equal(res["b200"], res["numpy"]) and np.abs(res["numpy"]).max() > 0, (name, ax)
        for name in ("first_order", "second_order", "third_order"):
            res = {}
            for backend in ("numpy", "b200"):
                obj = HorizontalSmoothing.factory(
                    f"{name}_1d{ax}", shape, 0.03, 0.24, 2, backend=backend,
                    backend_options=BackendOptions(), storage_options=StorageOptions())
                out = ta.zeros(backend, shape=shape)
                obj(ta.as_storage(backend, data=phi), out)
                res[backend] = np.array(to_numpy(out))
            assert np.array_equal(res["b200"], res["numpy"]), (name, ax)
    # the reference's class-less `diffusion` stencil (hyperdiffusion filter) on b200
    hbox = [int(v) for v in np.load(os.path.join(ROOT, "tests", "golden", "stencils_1d.npz"))["hyper_box"]]
    hfx = np.load(os.path.join(ROOT, "tests", "golden", "stencils_1d.npz"))
    hyper = ts.StencilCompiler("diffusion", "b200", backend_options=BackendOptions())
    hout = ta.zeros("b200", shape=hfx["hyper_phi"].shape)
    hyper(in_phi=ta.as_storage("b200", data=hfx["hyper_phi"]), out_phi=hout,
          alpha=float(hfx["hyper_alpha"]), origin=tuple(hbox[:3]), domain=tuple(hbox[3:]))
    assert np.array_equal(to_numpy(hout), hfx["hyper_out"]) and stub.count("tb200_hyperdiffusion") == 1
    # the reference's global `thomas` stencil compiled for b200 by the reference's own compiler
    fx1 = np.load(os.path.join(ROOT, "tests", "golden", "stencils_1d.npz"))
    tbox = [int(v) for v in fx1["thomas_box"]]
    thomas = ts.StencilCompiler("thomas", "b200", backend_options=BackendOptions())
    xs = ta.zeros("b200", shape=fx1["thomas_a"].shape)
    thomas(**{n: ta.as_storage("b200", data=fx1["thomas_" + n]) for n in "abcd"}, out=xs,
           origin=tuple(tbox[:3]), domain=tuple(tbox[3:]))
    assert np.array_equal(to_numpy(xs), fx1["thomas_x"]) and stub.count("tb200_thomas") == 1
    assert stub.count("tb200_diffusion_1d") == 8 and stub.count("tb200_smoothing_1d") == 12
    assert stub.count("tb200_diffusion") == 0 and stub.count("tb200_smoothing") == 0

# ---- the reference's one-dimensional boundary classes (Relaxed1DX / 1DY, Periodic1DX / 1DY:
# picked by its factory on grids with ny == 1 / nx == 1) run unchanged on backend b200 -- the
# `irelax` stencil through the C ABI, the slab copies as slice assignments on b200 storages
with stubbed_library(OracleStub) as stub:
    for hb_type, kw, nb in (("relaxed", {"nr": 5}, 2), ("periodic", {}, 2)):
        for nx, ny in ((17, 1), (1, 15)):
            nz = 4
            dom = gg._make_domain(nx, ny, nz, hb_type, nb, kw, topo=False)
            res = {}
            for backend in ("numpy", "b200"):
                hb1 = hbm.HorizontalBoundary.factory(
                    hb_type, dom.physical_grid, nb, backend=backend, backend_options=BackendOptions(),
                    storage_options=StorageOptions(), **kw)
                assert type(hb1).__name__.endswith("1DX" if ny == 1 else "1DY")
                shape = (hb1.ni + 1, hb1.nj + 1, nz + 1)
                r2 = np.random.default_rng(nx)
                fields = {n: r2.standard_normal(shape) for n in names}
                refs = {n: r2.standard_normal(shape) for n in names}
                conv = (lambda a: ta.as_storage("b200", data=a)) if backend == "b200" else (lambda a: a.copy())
                hb1.reference_state = {n: refload.DataArray(conv(v), attrs={"units": "1"})
                                       for n, v in refs.items()}
                got = {}
                for n in names:
                    f = conv(fields[n])
                    hb1.enforce_field(f, field_name=n, field_units="1")
                    hb1.set_outermost_layers_x(f, field_name=n, field_units="1")
                    hb1.set_outermost_layers_y(f, field_name=n, field_units="1")
                    got[n] = np.array(to_numpy(f))
                phys = r2.standard_normal((nx, ny, nz))
                got["numerical"] = np.array(to_numpy(hb1.get_numerical_field(conv(phys), field_name=names[0])))
                res[backend] = got
            for n in res["numpy"]:
                assert np.array_equal(res["numpy"][n], res["b200"][n]), (hb_type, nx, ny, n)
    assert stub.count("tb200_relax") == 8  # 2 relaxed boundaries x 4 fields

print("PLUGIN-OK", len(report["global"]), len(report["class_scoped"]), len(report["skipped"]))
for s in report["skipped"]:
    print("skipped:", s)
