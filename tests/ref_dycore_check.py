# -*- coding: utf-8 -*-
"""Run by tests/test_plugin_reference.py in a subprocess.  The UNMODIFIED reference dynamical core --
``IsentropicDynamicalCore.stage_array_call_dry`` and ``stage_array_call_moist``
(src/tasmania/isentropic/dynamics/dycore.py:L641-L843) with the reference's own Domain, Relaxed boundary, state builder, RK3WSSI prognostic, Rayleigh damper
HorizontalVelocity and WaterConstituent, all constructed with backend="b200" through the plugin -- is run for a full
RK3WS step against the oracle-backed C-ABI stub (tests/abi_oracle.py; storages on the host, every
kernel call carried out by the oracle), from the initial state of the golden fixtures
isen_{dry,moist}_rk3_5th.npz.  Two things are checked: (1) numerically, the first-stage outputs equal
the fixtures' -- which the reference's numpy backend wrote -- bit for bit, i.e. the unmodified
reference on backend b200 reproduces its own numpy backend through the plugin, the b200 stencil
wrappers and their marshalling; (2) the sequence of ABI calls it issues (kernel, canonical buffer
ids, scalars, boxes) is the one the b200 host mirror (tasmania_b200.isentropic.IsentropicDynamicalCore,
per-stencil path) issues from the same initial state.  Two representational differences are masked: the mirror
damps against the boundary's reference fields directly (the reference keeps copies) and stores
the Rayleigh coefficient at rank 1.  The mirror's extra launches (outermost layers and topography
factor as kernels instead of host-side slice assignments) are dropped from the comparison.
"""
import collections
import os
import sys
import types
from datetime import datetime, timedelta

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import refload  # noqa: E402

refload.install_framework()

import generate_golden as gg  # noqa: E402
import tasmania_b200 as tb  # noqa: E402
from tasmania_b200 import plugin  # noqa: E402
from tests.abi_oracle import OracleStub  # noqa: E402
from tests.abi_stub import stubbed_library  # noqa: E402

plugin.install(frame_relax=False, fused_stage=False)  # the reference-shaped per-stencil path on both sides

S, SU, SV = gg.S, gg.SU, gg.SV
U, V, MTG = "x_velocity_at_u_locations", "y_velocity_at_v_locations", "montgomery_potential"
P = gg.P
NB, NR = 3, 6
SCHEME, FLUX = "rk3ws_si", "fifth_order_upwind"
# the golden fixtures of these very configurations, written by the reference's numpy backend
# (tests/golden/generate_golden.py): their initial states are used, their first-stage outputs
# are what the b200-backend run must reproduce numerically
FIXTURES = {False: "isen_dry_rk3_5th", True: "isen_moist_rk3_5th"}
NX = NY = NZ = SHAPE = None  # set per case from the fixture


def set_case(moist):
    global NX, NY, NZ, SHAPE
    fx = np.load(os.path.join(ROOT, "tests", "golden", FIXTURES[moist] + ".npz"))
    NX, NY, NZ = (int(v) for v in fx["dims"][:3])
    assert tuple(int(v) for v in fx["dims"][3:5]) == (NB, NR) and int(fx["dims"][6]) == 4
    SHAPE = (NX + 1, NY + 1, NZ + 1)
    return fx
DT = timedelta(seconds=5)
OUTNAMES = (S, SU, U, SV, V)
QNAMES = (gg.MFWV, gg.MFCW, gg.MFPW)


def reference_trace(stub, moist, fx, fused=False, tendencies=None):
    """One RK3WS step of the reference's own dycore stage on backend b200; returns the ABI trace,
    the initial state as numpy arrays, the model-top pressure and the first-stage outputs (with
    ``fused``: through the plugin's fused-stage hook; the outputs of every stage)."""
    from tasmania.framework import allocators as ta
    from tasmania.framework.generic_functions import to_numpy

    DataArray = refload.DataArray
    dom = refload.load("tasmania.domain.domain")
    refload.load("tasmania.domain.subclasses.horizontal_boundaries.relaxed")
    refload.load("tasmania.domain.subclasses.topographies.gaussian")
    da = gg.da
    domain = dom.Domain(
        DataArray([-176, 176], dims="x", attrs={"units": "km"}), NX,
        DataArray([-176, 176], dims="y", attrs={"units": "km"}), NY,
        DataArray([400, 280], dims="z", attrs={"units": "K"}), NZ,
        horizontal_boundary_type="relaxed", nb=NB, horizontal_boundary_kwargs={"nr": NR},
        backend="b200", topography_type="gaussian",
        topography_kwargs={"time": timedelta(seconds=60), "max_height": da(0.5, "km"),
                           "width_x": da(50.0, "km"), "width_y": da(50.0, "km"), "smooth": False})
    grid = domain.numerical_grid
    st = refload.load("tasmania.isentropic.state")
    state = st.get_isentropic_state_from_brunt_vaisala_frequency(
        grid, datetime(2000, 1, 1), da(22.5, "m s^-1"), da(0.0, "m s^-1"), da(0.015, "s^-1"),
        moist=False, backend="b200", storage_shape=SHAPE)
    assert all(isinstance(v.data, tb.B200Array) for k, v in state.items() if k != "time")
    for n in list(state):  # the builder's own state equals the fixture's initial state
        if n != "time":
            assert np.array_equal(to_numpy(state[n].data), fx["init_" + n]), n
    if moist:  # the seeded water species of the moist fixture
        for n in QNAMES:
            state[n] = DataArray(ta.as_storage("b200", data=fx["init_" + n]), attrs={"units": "g g^-1"})
    hb = domain.horizontal_boundary
    hb.reference_state = state
    assert isinstance(hb._gamma, tb.B200Array)

    dyc = refload.load("tasmania.isentropic.dynamics.dycore")
    refload.load("tasmania.isentropic.dynamics.subclasses.prognostics.utils")
    refload.load("tasmania.isentropic.dynamics.subclasses.prognostics.rk3ws_si")
    refload.load("tasmania.isentropic.dynamics.subclasses.minimal_horizontal_fluxes.fifth_order_upwind")
    refload.load("tasmania.dwarfs.subclasses.vertical_dampers.rayleigh")
    prog = refload.load("tasmania.isentropic.dynamics.prognostic")
    vd = refload.load("tasmania.dwarfs.vertical_damping")
    dd = refload.load("tasmania.dwarfs.diagnostics")
    opts = refload.load("tasmania.framework.options")
    bo, so = opts.BackendOptions, opts.StorageOptions
    pt = float(to_numpy(state[P].data)[0, 0, 0])
    prognostic = prog.IsentropicPrognostic.factory(
        SCHEME, FLUX, domain, moist, backend="b200", backend_options=bo(), storage_shape=SHAPE,
        storage_options=so(), pt=da(pt, "Pa"), eps=0.5)
    damper = vd.VerticalDamping.factory("rayleigh", grid, 4, 5e-4, backend="b200", backend_options=bo(),
                                        storage_shape=SHAPE, storage_options=so())
    velocity = dd.HorizontalVelocity(grid, staggering=True, backend="b200", backend_options=bo(),
                                     storage_options=so())

    def zeros():
        return ta.zeros("b200", shape=SHAPE)

    outnames = OUTNAMES + (QNAMES if moist else ())
    water = {}
    if moist:
        water["_water_constituent"] = dd.WaterConstituent(
            grid, clipping=True, backend="b200", backend_options=bo(), storage_options=so())
        water.update({f"_{q}_{t}": zeros() for q in ("sqv", "sqc", "sqr") for t in ("now", "int", "new")})
    # the attributes stage_array_call_dry / _moist read from the dycore object
    me = types.SimpleNamespace(
        horizontal_boundary=hb, **water,
        # what the plugin's fused-stage hook reads in addition (attributes of the real object)
        backend="b200", grid=grid, storage_options=so(), _moist=moist,
        fast_tendency_component=None, fast_diagnostic_component=None,
        output_properties={k: {"units": state[k].attrs["units"]} for k in outnames},
        _damp=True, _damp_at_every_stage=True, stages=prognostic.stages, _prognostic=prognostic,
        _damper=damper, _velocity_components=velocity, _s_ref=zeros(), _su_ref=zeros(),
        _sv_ref=zeros(), _s_now=None, _su_now=None, _sv_now=None)
    cur = {k: state[k].data for k in (S, MTG, SU, U, SV, V) + (QNAMES if moist else ())}
    cur["time"] = state["time"]
    outs = [{k: zeros() for k in outnames} for _ in range(prognostic.stages)]
    stage_call = (dyc.IsentropicDynamicalCore.stage_array_call_moist if moist
                  else dyc.IsentropicDynamicalCore.stage_array_call_dry)
    if fused:  # looked up after the patch
        stage_call = (dyc.IsentropicDynamicalCore.stage_array_call_moist if moist
                      else dyc.IsentropicDynamicalCore.stage_array_call_dry)
    grid.update_topography(DT)
    stub.trace = []
    st_in = cur
    for stage in range(prognostic.stages):  # stage chaining of framework/dycore.py:L455-L458
        tnd = {k: ta.as_storage("b200", data=v) for k, v in (tendencies or {}).items()}
        stage_call(me, stage, st_in, tnd, DT, outs[stage])
        st_in = dict(outs[stage])
        st_in.setdefault(MTG, cur[MTG])
    trace, stub.trace = stub.trace, None
    stage0 = {k: to_numpy(v) for k, v in outs[0].items() if k != "time"}
    if fused:
        return trace, [{k: to_numpy(v) for k, v in o.items() if k != "time"} for o in outs], \
            [o["time"] for o in outs]
    return trace, {k: to_numpy(v.data) for k, v in state.items() if k != "time"}, pt, stage0


def mirror_trace(stub, np_state, pt, moist):
    from tasmania_b200.boundary import Relaxed
    from tasmania_b200.grid import Grid, Topography, gaussian_profile
    from tasmania_b200.isentropic import IsentropicDynamicalCore

    os.environ["TB200_RELAX"] = "full"  # the reference-shaped boundary path: one irelax per field
    x, y = np.linspace(-176.0, 176.0, NX), np.linspace(-176.0, 176.0, NY)
    grid = Grid((-176.0, 176.0), NX, (-176.0, 176.0), NY, (400.0, 280.0), NZ, units_to_m=1e3,
                topography=Topography(gaussian_profile(x, y, 500.0, 50.0, 50.0), timedelta(seconds=60)))
    hb = Relaxed(NX, NY, NZ, NB, nr=NR)
    state = {k: tb.as_storage(v) for k, v in np_state.items()}
    state["time"] = datetime(2000, 1, 1)
    hb.reference_state = state
    dycore = IsentropicDynamicalCore(
        grid, hb, moist=moist, time_integration_scheme=SCHEME, horizontal_flux_scheme=FLUX,
        time_integration_properties={"pt": pt, "eps": 0.5}, damp=True, damp_depth=4, damp_max=5e-4,
        fused=False)
    dycore.update_topography(DT)
    stub.trace = []
    dycore(state, {}, DT)
    trace, stub.trace = stub.trace, None
    return trace


# argument positions whose buffer identity is a representational choice (see the module docstring)
MASK = {"tb200_damping": {2, 3}}
DROPPED = ("tb200_set_outermost_layers", "tb200_elementwise")


def reduce(trace):
    """Kernel name, scalars, boxes and canonical buffer ids (numbered by first appearance)."""
    ids, out = {}, []

    def is_field(x):
        return isinstance(x, tuple) and len(x) == 3 and isinstance(x[1], tuple) and isinstance(x[0], int)

    for name, desc in trace:
        if name in DROPPED:
            continue
        row = [name]
        for pos, x in enumerate(desc):
            if pos in MASK.get(name, ()):
                row.append("masked")
            elif is_field(x):
                row.append(("field", ids.setdefault(x[0], len(ids))))
            elif isinstance(x, tuple) and x and isinstance(x[0], tuple):
                row.append(tuple(("field", ids.setdefault(f[0], len(ids))) if f else None for f in x))
            else:
                row.append(x)
        out.append(tuple(row))
    return out


EXPECTED = {
    # per stage: K1, irelax(s), montgomery, K2, five irelax, three dampings, two velocity diagnoses
    False: {"tb200_relax": 18, "tb200_damping": 9, "tb200_velocity": 6, "tb200_step_forward_euler": 3,
            "tb200_montgomery": 3, "tb200_step_forward_euler_momentum": 3},
    # + sq = s q of the three species before, q = sq / s after, and their three relaxations
    True: {"tb200_relax": 27, "tb200_damping": 9, "tb200_velocity": 6, "tb200_step_forward_euler": 3,
           "tb200_montgomery": 3, "tb200_step_forward_euler_momentum": 3, "tb200_density": 9,
           "tb200_mass_fraction": 9},
}
done = []
for moist_case in (False, True):
    fixture = set_case(moist_case)
    # the oracle carries the kernels out on the host buffers: the run is numerical
    with stubbed_library(OracleStub) as the_stub:
        ref_trace, initial, p_top, first_stage = reference_trace(the_stub, moist_case, fixture)
        counts = collections.Counter(n for n, _ in ref_trace)
        assert counts == EXPECTED[moist_case], counts
        mir_trace = mirror_trace(the_stub, initial, p_top, moist_case)
    # the unmodified reference dycore on backend b200 reproduces its own numpy backend
    for name, got in first_stage.items():
        assert np.array_equal(got, fixture["stage0_" + name]), (moist_case, name)
    a, b = reduce(ref_trace), reduce(mir_trace)
    assert len(a) == len(b) == sum(EXPECTED[moist_case].values()), (len(a), len(b))
    for n, (p, q) in enumerate(zip(a, b)):
        assert p == q, (moist_case, n, p, q)
    done.append(len(a))

# ---- the fused stage behind the reference's own class (plugin.install(fused_stage=True)): the
# unmodified reference objects, stage_array_call_dry patched for backend b200, (1) issue one
# tb200_isentropic_stage_dry call per stage with the lazy-velocity flags, (2) reproduce the
# reference's numpy backend on the first stage (s, su, sv: the fixture) and (3) end the step with
# exactly the fields the per-stencil reference run ends with
fixture = set_case(False)
with stubbed_library(OracleStub) as the_stub:
    assert plugin._fused_stage_dry() == "IsentropicDynamicalCore.stage_array_call_dry"
    assert plugin._fused_stage_dry() is None  # idempotent
    fused_trace, fused_outs, fused_times = reference_trace(the_stub, False, fixture, fused=True)
names = [n for n, _ in fused_trace]
assert names.count("tb200_isentropic_stage_dry") == 3, collections.Counter(names)
assert not any(n in names for n in ("tb200_step_forward_euler", "tb200_velocity", "tb200_damping", "tb200_relax"))
flags = [(dict(zip([f for f, _ in tb.lib.StageCfg._fields_], d[0]))["derive_uv_in"],
          dict(zip([f for f, _ in tb.lib.StageCfg._fields_], d[0]))["skip_uv_out"])
         for n, d in fused_trace if n == "tb200_isentropic_stage_dry"]
assert flags == [(0, 1), (1, 1), (1, 1)], flags  # no stage writes u, v ...
assert names.count("tb200_velocity_components") == 1 and names[-1] == "tb200_velocity_components"  # ... one pass does
for name in (S, SU, SV):
    assert np.array_equal(fused_outs[0][name], fixture["stage0_" + name]), name
# never written by an intermediate stage (stub poison over the box a kernel would write)
assert np.isnan(fused_outs[0][U]).any() and np.isnan(fused_outs[1][V]).any()
# the per-stencil reference run of the same step, all stages, for the final fields
import importlib  # noqa: E402

dyc_mod = importlib.import_module("tasmania.isentropic.dynamics.dycore")
patched = dyc_mod.IsentropicDynamicalCore.stage_array_call_dry
dyc_mod.IsentropicDynamicalCore.stage_array_call_dry = patched.__wrapped_original__
try:
    with stubbed_library(OracleStub) as the_stub:
        _, plain_outs, plain_times = reference_trace(the_stub, False, fixture, fused=True)
finally:
    dyc_mod.IsentropicDynamicalCore.stage_array_call_dry = patched
for name in OUTNAMES:
    assert np.array_equal(fused_outs[2][name], plain_outs[2][name]), name
assert fused_times == plain_times
# ---- the same for the moist stage: stage_array_call_moist patched -> one
# tb200_isentropic_stage_moist per stage, no density / mass_fraction / relax launches, and the step
# ends with exactly the fields of the per-stencil reference run (water constituents included)
fixture = set_case(True)
with stubbed_library(OracleStub) as the_stub:
    assert plugin._fused_stage_moist() == "IsentropicDynamicalCore.stage_array_call_moist"
    assert plugin._fused_stage_moist() is None  # idempotent
    fused_trace, fused_outs, fused_times = reference_trace(the_stub, True, fixture, fused=True)
names = [n for n, _ in fused_trace]
assert names.count("tb200_isentropic_stage_moist") == 3 and names[-1] == "tb200_velocity_components", \
    collections.Counter(names)
assert not any(n in names for n in ("tb200_step_forward_euler", "tb200_density", "tb200_mass_fraction",
                                    "tb200_damping", "tb200_relax", "tb200_isentropic_stage_dry"))
for name in (S, SU, SV) + QNAMES:
    assert np.array_equal(fused_outs[0][name], fixture["stage0_" + name]), name
patched = dyc_mod.IsentropicDynamicalCore.stage_array_call_moist
dyc_mod.IsentropicDynamicalCore.stage_array_call_moist = patched.__wrapped_original__
try:
    with stubbed_library(OracleStub) as the_stub:
        _, plain_outs, plain_times = reference_trace(the_stub, True, fixture, fused=True)
finally:
    dyc_mod.IsentropicDynamicalCore.stage_array_call_moist = patched
for name in OUTNAMES + QNAMES:
    assert np.array_equal(fused_outs[2][name], plain_outs[2][name]), name
assert fused_times == plain_times
# ---- slow tendencies of su, sv (a subset: s gets a field of zeros) through the fused DRY hook: one
# tb200_isentropic_stage_dry per stage carrying the three tendency pointers, same final fields as the
# per-stencil reference run with the same tendencies
fixture = set_case(False)
rng = np.random.default_rng(17)
slow = {SU: 1e-2 * rng.standard_normal(SHAPE), SV: 1e-2 * rng.standard_normal(SHAPE)}
with stubbed_library(OracleStub) as the_stub:
    fused_trace, fused_outs, _ = reference_trace(the_stub, False, fixture, fused=True, tendencies=slow)
names = [n for n, _ in fused_trace]
assert names.count("tb200_isentropic_stage_dry") == 3 and "tb200_step_forward_euler" not in names, \
    collections.Counter(names)
fields = [f for f, _ in tb.lib.StageCfg._fields_]
for n, d in fused_trace:
    if n == "tb200_isentropic_stage_dry":
        cfg = dict(zip(fields, d[0]))
        assert cfg["s_tnd"] is not None and cfg["su_tnd"] is not None and cfg["sv_tnd"] is not None
patched = dyc_mod.IsentropicDynamicalCore.stage_array_call_dry
dyc_mod.IsentropicDynamicalCore.stage_array_call_dry = patched.__wrapped_original__
try:
    with stubbed_library(OracleStub) as the_stub:
        plain_trace, plain_outs, _ = reference_trace(the_stub, False, fixture, fused=True, tendencies=slow)
finally:
    dyc_mod.IsentropicDynamicalCore.stage_array_call_dry = patched
assert "tb200_step_forward_euler" in [n for n, _ in plain_trace]
for name in OUTNAMES:
    assert np.array_equal(fused_outs[2][name], plain_outs[2][name]), name
assert not np.array_equal(fused_outs[2][SU][3:-4, 3:-4, :-1], fixture["stage0_" + SU][3:-4, 3:-4, :-1])
print("REF-DYCORE-OK", *done, "FUSED-HOOK-OK", "FUSED-MOIST-HOOK-OK", "FUSED-TENDENCIES-HOOK-OK")
