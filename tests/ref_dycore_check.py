# -*- coding: utf-8 -*-
"""Run by tests/test_plugin_reference.py in a subprocess.  The UNMODIFIED reference dynamical core --
``IsentropicDynamicalCore.stage_array_call_dry`` (src/tasmania/isentropic/dynamics/dycore.py:L641-L721)
with the reference's own Domain, Relaxed boundary, state builder, RK3WSSI prognostic, Rayleigh damper
and HorizontalVelocity, all constructed with backend="b200" through the plugin -- is run for a full
RK3WS step against the recording C-ABI stub (tests/abi_stub.py; storages on the host), and the
sequence of ABI calls it issues (kernel, canonical buffer ids, scalars, boxes) is compared with the
one the b200 host mirror (tasmania_b200.isentropic.IsentropicDynamicalCore, per-stencil path)
issues from the same initial state.  Two representational differences are masked: the mirror
damps against the boundary's reference fields directly (the reference keeps copies) and stores
the Rayleigh coefficient at rank 1.  The mirror's extra launches (outermost layers and topography
factor as kernels instead of host-side slice assignments) are dropped from the comparison.
"""
import collections
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"tests","golden"))
import numpy as np
from datetime import datetime, timedelta
import refload
refload.install_framework()
import generate_golden as gg
from tasmania_b200 import plugin, storage
from tests.abi_stub import stubbed_library, canonical
plugin.install()
S,SU,SV,U,V,MTG = gg.S, gg.SU, gg.SV, "x_velocity_at_u_locations","y_velocity_at_v_locations","montgomery_potential"
P,EXN,H = gg.P, gg.EXN, gg.H
nx,ny,nz,nb,nr = 25,21,8,3,6
scheme, flux = "rk3ws_si", "fifth_order_upwind"
with stubbed_library() as stub:
    dom = refload.load("tasmania.domain.domain")
    for m in ("relaxed",): refload.load("tasmania.domain.subclasses.horizontal_boundaries."+m)
    refload.load("tasmania.domain.subclasses.topographies.gaussian")
    DataArray = refload.DataArray
    d = dom.Domain(DataArray([-176,176], dims="x", attrs={"units":"km"}), nx,
                   DataArray([-176,176], dims="y", attrs={"units":"km"}), ny,
                   DataArray([400,280], dims="z", attrs={"units":"K"}), nz,
                   horizontal_boundary_type="relaxed", nb=nb, horizontal_boundary_kwargs={"nr":nr},
                   backend="b200", topography_type="gaussian",
                   topography_kwargs={"time": timedelta(seconds=60), "max_height": gg.da(0.5,"km"),
                                      "width_x": gg.da(50.0,"km"), "width_y": gg.da(50.0,"km"), "smooth": False})
    g = d.numerical_grid
    st = refload.load("tasmania.isentropic.state")
    shape=(nx+1,ny+1,nz+1)
    state = st.get_isentropic_state_from_brunt_vaisala_frequency(
        g, datetime(2000,1,1), gg.da(22.5,"m s^-1"), gg.da(0.0,"m s^-1"), gg.da(0.015,"s^-1"),
        moist=False, backend="b200", storage_shape=shape)
    assert all(type(v.data).__name__ == "B200Array" for k, v in state.items() if k != "time")
    hb = d.horizontal_boundary
    hb.reference_state = state
    assert type(hb._gamma).__name__ == "B200Array"
    dyc = refload.load("tasmania.isentropic.dynamics.dycore")
    refload.load("tasmania.isentropic.dynamics.subclasses.prognostics.utils")
    refload.load("tasmania.isentropic.dynamics.subclasses.prognostics.rk3ws_si")
    refload.load("tasmania.isentropic.dynamics.subclasses.minimal_horizontal_fluxes.fifth_order_upwind")
    refload.load("tasmania.dwarfs.subclasses.vertical_dampers.rayleigh")
    prog = refload.load("tasmania.isentropic.dynamics.prognostic")
    vd = refload.load("tasmania.dwarfs.vertical_damping")
    dd = refload.load("tasmania.dwarfs.diagnostics")
    opts = refload.load("tasmania.framework.options")
    from tasmania.framework.generic_functions import to_numpy
    from tasmania.framework import allocators as ta
    pt = float(to_numpy(state[P].data)[0,0,0])
    bo, so = opts.BackendOptions, opts.StorageOptions
    P_ = prog.IsentropicPrognostic.factory(scheme, flux, d, False, backend="b200", backend_options=bo(),
                                           storage_shape=shape, storage_options=so(), pt=gg.da(pt,"Pa"), eps=0.5)
    damper = vd.VerticalDamping.factory("rayleigh", g, 4, 5e-4, backend="b200", backend_options=bo(),
                                        storage_shape=shape, storage_options=so())
    vel = dd.HorizontalVelocity(g, staggering=True, backend="b200", backend_options=bo(), storage_options=so())
    z = lambda: ta.zeros("b200", shape=shape)
    outnames=(S,SU,U,SV,V)
    fake = types.SimpleNamespace(horizontal_boundary=hb,
        output_properties={k: {"units": state[k].attrs["units"]} for k in outnames},
        _damp=True, _damp_at_every_stage=True, stages=P_.stages, _prognostic=P_, _damper=damper,
        _velocity_components=vel, _s_ref=z(), _su_ref=z(), _sv_ref=z(), _s_now=None, _su_now=None, _sv_now=None)
    cur = {k: state[k].data for k in (S,MTG,SU,U,SV,V)}
    cur["time"] = state["time"]
    outs = [{k: z() for k in outnames} for _ in range(P_.stages)]
    dt = timedelta(seconds=5)
    g.update_topography(dt)
    stub.trace = []
    st_in = cur
    for stage in range(P_.stages):
        dyc.IsentropicDynamicalCore.stage_array_call_dry(fake, stage, st_in, {}, dt, outs[stage])
        st_in = dict(outs[stage]); st_in.setdefault(MTG, cur[MTG])
    tr = stub.trace; stub.trace=None
    counts = collections.Counter(n for n, _ in tr)
    # 13 reference passes per stage + the second relaxation of s
    assert counts == {"tb200_relax": 18, "tb200_damping": 9, "tb200_velocity": 6,
                      "tb200_step_forward_euler": 3, "tb200_montgomery": 3,
                      "tb200_step_forward_euler_momentum": 3}, counts
    # ---- the mirror, unfused, same initial state
    os.environ["TB200_RELAX"] = "full"
    from tasmania_b200.boundary import Relaxed
    from tasmania_b200.grid import Grid, Topography, gaussian_profile
    from tasmania_b200.isentropic import IsentropicDynamicalCore
    import tasmania_b200 as tb
    x, y = np.linspace(-176.0,176.0,nx), np.linspace(-176.0,176.0,ny)
    grid = Grid((-176.0,176.0), nx, (-176.0,176.0), ny, (400.0,280.0), nz, units_to_m=1e3,
                topography=Topography(gaussian_profile(x,y,500.0,50.0,50.0), timedelta(seconds=60)))
    mhb = Relaxed(nx,ny,nz,nb,nr=nr)
    mstate = {k: tb.as_storage(to_numpy(v.data)) for k,v in state.items() if k!="time"}
    mstate["time"] = datetime(2000,1,1)
    mhb.reference_state = mstate
    mdyc = IsentropicDynamicalCore(grid, mhb, time_integration_scheme=scheme, horizontal_flux_scheme=flux,
                                   time_integration_properties={"pt": pt, "eps": 0.5}, damp=True, damp_depth=4,
                                   damp_max=5e-4, fused=False)
    mdyc.update_topography(dt)
    stub.trace = []
    mdyc(mstate, {}, dt)
    mtr = stub.trace; stub.trace=None

MASK = {'tb200_damping': {2, 3}}
def reduce(trace):
    ids = {}
    out = []
    for name, desc in trace:
        if name in ("tb200_set_outermost_layers", "tb200_elementwise"):
            continue
        row = [name]
        for pos, x in enumerate(desc):
            if pos in MASK.get(name, ()):
                row.append("masked")
                continue
            if isinstance(x, tuple) and len(x) == 3 and isinstance(x[1], tuple) and isinstance(x[0], int):
                row.append(("field", ids.setdefault(x[0], len(ids))))
            elif isinstance(x, tuple) and x and isinstance(x[0], tuple):
                row.append(tuple(("field", ids.setdefault(f[0], len(ids))) if f else None for f in x))
            else:
                row.append(x)
        out.append(tuple(row))
    return out
a, b = reduce(tr), reduce(mtr)
assert len(a) == len(b) == 42, (len(a), len(b))
for n, (p, q) in enumerate(zip(a, b)):
    assert p == q, (n, p, q)
print("REF-DYCORE-OK", len(a))
