# -*- coding: utf-8 -*-
"""Pin of the coupling layer on the REFERENCE ITSELF, executed in place from /root/reference
(tests/golden/refload.py; skipped where the reference tree is absent, e.g. on the GPU box).

The reference's couplers cannot be instantiated here -- their constructors and ``__call__``
wrappers live in the private sympl fork -- but the methods that hold the coupling *logic* are
tasmania's own and run unmodified when called unbound on a stand-in ``self``:

  ForwardEuler._call / RK2._call / RK3WS._call   framework/subclasses/tendency_steppers/*.py
  DataArrayDictOperator.fma / update_swap         utils/xarrayx.py
  ConcurrentCoupling._call_serial                 framework/concurrent_coupling.py:L314-L374
  StaticOperator.get_overwrite_tendencies         framework/concurrent_coupling_utils.py:L72-L83
  From*To*.array_call                             framework/promoter.py:L161-L176, L290-L305
  SequentialUpdateSplitting.__call__              framework/sequential_update_splitting.py:L161-L194

Their results must equal, bit for bit, the oracle's straight-line ``tendency_step`` and the b200
host mirrors of tasmania_b200/coupling.py (run on host storages through tests/abi_stub.py).
The only sympl-side behaviour assumed is ``StaticComponentOperator.get_properties(component)``
= the component's ``<name>_properties`` attribute.
"""
import importlib
import types
from datetime import datetime, timedelta

import numpy as np
import pytest

from oracle import moist_model as mm
from tests.abi_stub import stubbed_library
from tests.golden import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference tree not mounted")

THETA, W = "air_potential_temperature", "tendency_of_air_potential_temperature"
DIMS = ("x", "y", "z")


def ref(mod):
    refload.install_framework()
    return importlib.import_module(mod)


def da(a, units="m"):
    return refload.DataArray(np.array(a, dtype=float), None, DIMS, None, {"units": units})


class Properties:
    """sympl's StaticComponentOperator as far as the coupling code uses it."""

    def __init__(self, name):
        self.name = name

    @classmethod
    def factory(cls, name):
        return cls(name)

    def get_properties(self, component):
        return getattr(component, self.name, {})


# ------------------------------------------------------------------ tendency steppers
@pytest.mark.parametrize("scheme,modname,clsname", [("forward_euler", "forward_euler", "ForwardEuler"),
                                                    ("rk2", "rk2", "RK2"), ("rk3ws", "rk3ws", "RK3WS")])
def test_reference_stepper_logic_equals_oracle_and_b200(scheme, modname, clsname):
    import tasmania_b200 as tb
    from tasmania_b200.coupling import TendencyStepper
    from tests.test_coupling_host import Decay

    cls = getattr(ref("tasmania.framework.subclasses.tendency_steppers." + modname), clsname)
    op = ref("tasmania.utils.xarrayx").DataArrayDictOperator(backend="numpy")
    rng = np.random.default_rng(1)
    y0, z0 = rng.standard_normal((4, 3, 2)), rng.standard_normal((4, 3, 2))
    dt = timedelta(seconds=0.3)

    class Increment:  # sympl's get_increment: the tendencies of the wrapped component
        def get_increment(self, st, timestep, out_increment=None, out_diagnostics=None):
            inc = {"y": da(-1.0 * st["y"].data + -0.5 * st["y"].data), "z": da(-1.0 * st["z"].data),
                   "time": st["time"]}
            return inc, {"seen": da(st["y"].data)}

    state = {"y": da(y0), "z": da(z0), "other": da(y0), "time": datetime(2000, 1, 1)}
    fake = types.SimpleNamespace(
        _stepper_operator=Increment(), _dict_op=op, _enforce_hb=False, _increment=None, _diagnostics=None,
        output_properties={"y": {"units": "m", "dims": DIMS}, "z": {"units": "m", "dims": DIMS}})
    out_d, out_s = cls._call(fake, state, dt, {}, {"y": da(np.zeros_like(y0)), "z": da(np.zeros_like(y0))})
    assert set(out_s) == {"y", "z", "time"} and out_s["time"] == datetime(2000, 1, 1) + dt
    np.testing.assert_array_equal(out_d["seen"].data, y0)  # diagnostics of the first stage

    def fn(st):
        return {"y": -1.0 * st["y"] + -0.5 * st["y"], "z": -1.0 * st["z"]}, {}

    _, want = mm.tendency_step(scheme, {"y": y0, "z": z0, "other": y0}, fn, dt.total_seconds())
    with stubbed_library():
        dstate = {"y": tb.as_storage(y0), "z": tb.as_storage(z0), "other": tb.as_storage(y0),
                  "time": datetime(2000, 1, 1)}
        _, dout = TendencyStepper.factory(scheme, Decay(1.0, ("y", "z")), Decay(0.5, ("y",)))(dstate, dt)
        got = {n: tb.to_numpy(dout[n]) for n in ("y", "z")}
    for n in ("y", "z"):
        np.testing.assert_array_equal(out_s[n].data, want[n], err_msg=f"reference vs oracle, {n}")
        np.testing.assert_array_equal(got[n], out_s[n].data, err_msg=f"b200 vs reference, {n}")


# ------------------------------------------------------------------ concurrent coupling + promoters
def _promoter(cls, grid, **props):
    """An instance of a reference promoter class without its sympl-side constructor; ``__call__``
    goes straight to the reference's own ``array_call`` on the raw arrays."""
    copy_numpy = ref("tasmania.framework.subclasses.stencil_definitions.copy").copy_numpy

    class Promoter(cls):
        def __init__(self):
            pass

        def __call__(self, arrays, *, out=None):
            self.array_call({n: v.data for n, v in arrays.items() if n != "time"},
                            {n: v.data for n, v in out.items() if n != "time"})
            return out

    p = Promoter()
    p.__dict__.update(props)
    p._stencil_copy = lambda src, dst, origin, domain, **kw: copy_numpy(src, dst, origin=origin, domain=domain)
    p._backend_options = types.SimpleNamespace(exec_info=None, validate_args=False)
    Promoter.grid = grid
    Promoter.backend_options = property(lambda self: self._backend_options)
    return p


def test_reference_serial_coupling_with_promoters_equals_b200():
    import tasmania_b200 as tb
    from tasmania_b200 import coupling as bc
    from tasmania_b200.grid import Grid

    cc = ref("tasmania.framework.concurrent_coupling")
    ccu = ref("tasmania.framework.concurrent_coupling_utils")
    iu = ref("tasmania.isentropic.utils")
    cc.StaticComponentOperator = Properties
    ccu.StaticOperator.tendency_operator = Properties("tendency_properties")

    nx, ny, nz = 3, 2, 1
    axis = lambda n: types.SimpleNamespace(dims=(n,))  # noqa: E731
    grid = types.SimpleNamespace(nx=nx, ny=ny, nz=nz, x=axis("x"), y=axis("y"), z=axis("z"))
    w0 = np.arange(24.0).reshape(4, 3, 2) + 1.0
    prop = {"dims": DIMS, "units": "K s^-1"}

    class Heating(cc.TendencyComponent):  # adds 2 K/s, honouring the overwrite flag it is given
        tendency_properties = {THETA: prop}
        diagnostic_properties = {}
        seen = None

        def __call__(self, state, out_tendencies=None, out_diagnostics=None, overwrite_tendencies=None):
            Heating.seen = dict(overwrite_tendencies)
            t = out_tendencies[THETA].data
            t[...] = 2.0 if overwrite_tendencies[THETA] else t + 2.0

    d2t = _promoter(iu.AirPotentialTemperatureToTendency, grid, tendency_properties={THETA: prop},
                    diagnostic_properties={})
    t2d = _promoter(iu.AirPotentialTemperatureToDiagnostic, grid, diagnostic_properties={W: prop},
                    tendency_properties={})
    fake = types.SimpleNamespace(
        components=(d2t, Heating(), t2d), execution_policy="serial",
        allowed_diagnostic_type=cc.ConcurrentCoupling.allowed_diagnostic_type,
        allowed_tendency_type=cc.ConcurrentCoupling.allowed_tendency_type)
    fake.overwrite_tendencies = ccu.StaticOperator.get_overwrite_tendencies(fake)
    assert fake.overwrite_tendencies == [{THETA: True}, {THETA: False}, {}]
    state = {W: da(w0, "K s^-1"), "time": datetime(2000, 1, 1)}
    tnd, diag = {THETA: da(np.zeros_like(w0), "K s^-1")}, {W: da(np.zeros_like(w0), "K s^-1")}
    cc.ConcurrentCoupling._call_serial(fake, state, timedelta(seconds=1), tnd, diag, {})
    assert Heating.seen == {THETA: False}
    want = np.zeros_like(w0)
    want[:nx, :ny, :nz] = w0[:nx, :ny, :nz] + 2.0
    np.testing.assert_array_equal(diag[W].data, want)

    class B200Heating:
        kind, diagnostic_names, tendency_names = "tendency", (), (THETA,)

        def array_call(self, state, out_tendencies, out_diagnostics, overwrite_tendencies):
            assert overwrite_tendencies == Heating.seen
            out_tendencies[THETA].t.add_(2.0)

    with stubbed_library():
        g = Grid((0.0, 1.0), nx, (0.0, 1.0), ny, (300.0, 280.0), nz)
        coupler = bc.ConcurrentCoupling(bc.AirPotentialTemperatureToTendency(g), B200Heating(),
                                        bc.AirPotentialTemperatureToDiagnostic(g))
        assert coupler.overwrite_tendencies == fake.overwrite_tendencies
        dstate = {W: tb.as_storage(w0), "air_isentropic_density": tb.as_storage(w0)}
        _, ddiag = coupler(dstate, timedelta(seconds=1))
        np.testing.assert_array_equal(tb.to_numpy(ddiag[W]), diag[W].data)


# ------------------------------------------------------------------ sequential-update splitting
def test_reference_sequential_update_splitting_equals_b200():
    import tasmania_b200 as tb
    from tasmania_b200 import coupling as bc
    from tests.test_coupling_host import Decay

    sus = ref("tasmania.framework.sequential_update_splitting")
    rk2 = ref("tasmania.framework.subclasses.tendency_steppers.rk2").RK2
    op = ref("tasmania.utils.xarrayx").DataArrayDictOperator(backend="numpy")
    y0 = np.random.default_rng(2).standard_normal((4, 3, 2))
    dt = timedelta(seconds=0.25)

    class Doubler(sus.DiagnosticComponent):  # a diagnostic component overwriting y
        def __call__(self, state, out=None):
            out = out if out is not None else {"y": da(np.zeros_like(y0))}
            out["y"].data[...] = 2.0 * state["y"].data
            return out

    class Stepper:  # the sympl-side wrapper of a TendencyStepper: allocate, then the reference's _call
        _dict_op, _enforce_hb, _increment, _diagnostics = op, False, None, None
        output_properties = {"y": {"units": "m", "dims": DIMS}}

        class _stepper_operator:
            @staticmethod
            def get_increment(st, timestep, out_increment=None, out_diagnostics=None):
                return {"y": da(-0.7 * st["y"].data), "time": st["time"]}, {"seen": da(st["y"].data)}

        def __call__(self, state, timestep, out_diagnostics=None, out_state=None):
            out_state = out_state if out_state is not None else {"y": da(np.zeros_like(y0))}
            out_diagnostics = out_diagnostics if out_diagnostics is not None else {}
            return rk2._call(self, state, timestep, out_diagnostics, out_state)

    fake = types.SimpleNamespace(
        _component_list=[Doubler(), Stepper()], _substeps=[1, 1], _out_diagnostics=[None, None],
        _out_state=[None, None], _dict_op=op,
        allowed_diagnostic_type=sus.SequentialUpdateSplitting.allowed_diagnostic_type)
    state = {"y": da(y0), "time": datetime(2000, 1, 1)}
    for _ in range(3):
        sus.SequentialUpdateSplitting.__call__(fake, state, dt)
    assert state["time"] == datetime(2000, 1, 1) + 3 * dt

    class B200Doubler:
        kind, tendency_names, diagnostic_names = "diagnostic", (), ("y",)

        def diagnostic_shape(self, name):
            return y0.shape

        def zeros(self, *, shape):
            return tb.zeros(shape)

        def array_call(self, state, out):
            out["y"].t.copy_(2.0 * state["y"].t)

    class B200Decay(Decay):
        def array_call(self, state, out_tendencies, out_diagnostics, overwrite_tendencies):
            out_tendencies["y"].t.copy_(-0.7 * state["y"].t)
            out_diagnostics["seen"].t.copy_(state["y"].t)

    with stubbed_library():
        dstate = {"y": tb.as_storage(y0), "time": datetime(2000, 1, 1)}
        dsus = bc.SequentialUpdateSplitting(bc.TimeIntegrationOptions(B200Doubler()),
                                            bc.TimeIntegrationOptions(B200Decay(0.7), scheme="rk2"))
        for _ in range(3):
            dsus(dstate, dt)
        np.testing.assert_array_equal(tb.to_numpy(dstate["y"]), state["y"].data)
        np.testing.assert_array_equal(tb.to_numpy(dstate["seen"]), state["seen"].data)
        assert dstate["time"] == state["time"]


# ------------------------------------------------------------------ moist initial state
def test_moist_initial_state_equals_reference():
    """tasmania_b200.grid.isentropic_state_from_brunt_vaisala(moist=True) restates host-side set-up
    (src/tasmania/isentropic/state.py:L126-L391, utils/meteo.py:L192-L274); here against the
    reference's own function run in place: every field bit for bit."""
    from tasmania_b200.grid import Grid, Topography, gaussian_profile, isentropic_state_from_brunt_vaisala
    from tests.golden import generate_golden as gg

    nx, ny, nz = 21, 19, 10
    d = gg._make_domain(nx, ny, nz, "relaxed", 3, {"nr": 6}, topo_time=60.0)
    st = ref("tasmania.isentropic.state")
    shape = (nx + 1, ny + 1, nz + 1)
    want = st.get_isentropic_state_from_brunt_vaisala_frequency(
        d.numerical_grid, datetime(2000, 1, 1), gg.da(22.5, "m s^-1"), gg.da(0.0, "m s^-1"),
        gg.da(0.015, "s^-1"), moist=True, precipitation=True, relative_humidity=0.95, backend="numpy",
        storage_shape=shape)
    x, y = np.linspace(-176.0, 176.0, nx), np.linspace(-176.0, 176.0, ny)
    grid = Grid((-176.0, 176.0), nx, (-176.0, 176.0), ny, (400.0, 280.0), nz, units_to_m=1e3,
                topography=Topography(gaussian_profile(x, y, 500.0, 50.0, 50.0), timedelta(seconds=60)))
    got = isentropic_state_from_brunt_vaisala(grid, 22.5, 0.0, 0.015, moist=True, precipitation=True,
                                              relative_humidity=0.95)
    assert set(got) == set(want) - {"time"}
    for n, v in got.items():
        np.testing.assert_array_equal(v, np.asarray(want[n].data), err_msg=n)
    assert float(got["mass_fraction_of_water_vapor_in_air"].max()) > 1e-3


def test_smoothing_with_boundary_layers_wider_than_the_stencil_equals_reference():
    """The benchmark smooths with nb = 3 (the boundary's) under a second-order (5-point) smoother
    (horizontal_smoothing.py:L113-L126 with second_order.py:L48); the reference's own smoother,
    run in place, against the call the oracle's moist model makes."""
    from oracle import dwarfs

    hsm = ref("tasmania.dwarfs.horizontal_smoothing")
    ref("tasmania.dwarfs.subclasses.horizontal_smoothers.second_order")
    opts = ref("tasmania.framework.options")
    shape, nb = (19, 17, 7), 3
    phi = np.random.default_rng(4).standard_normal(shape)
    obj = hsm.HorizontalSmoothing.factory(
        "second_order", shape, 1.0, 1.0, 0, nb, backend="numpy",
        backend_options=opts.BackendOptions(), storage_options=opts.StorageOptions())
    want = np.zeros(shape)
    obj(phi, want)
    gamma = np.zeros(shape)
    gamma[...] = dwarfs.vertical_profile(1.0, 1.0, 0, shape[2])[None, None, :]
    np.testing.assert_array_equal(gamma, np.asarray(obj._gamma))
    got = np.zeros(shape)
    sx, sy, sz = shape
    dwarfs.smoothing(2, phi, gamma, got, (nb, nb, 0), (sx - 2 * nb, sy - 2 * nb, sz))
    for o, d in (((0, 0, 0), (nb, sy, sz)), ((sx - nb, 0, 0), (nb, sy, sz)),
                 ((nb, 0, 0), (sx - 2 * nb, nb, sz)), ((nb, sy - nb, 0), (sx - 2 * nb, nb, sz))):
        dwarfs.copy(phi, got, o, d)
    np.testing.assert_array_equal(got, want)
    assert not np.array_equal(want[nb:-nb, nb:-nb], phi[nb:-nb, nb:-nb])
    np.testing.assert_array_equal(want[:nb], phi[:nb])
