# -*- coding: utf-8 -*-
"""The ``b200`` stencil definitions: one per reference stencil name on the hot path
(SURVEY.md section 8a), with the reference's keyword names.  Each is a thin marshalling layer
over one C-ABI entry point of ``libtasmania_b200.so``; the arithmetic lives in the CUDA
kernels.  ``externals`` is the snapshot of ``BackendOptions.externals`` taken at
``compile_stencil`` time and selects the kernel variant (flux scheme, ``moist``, constants).
"""
from __future__ import annotations

import ctypes as C

from tasmania_b200 import lib
from tasmania_b200.framework import stencil_definition, subroutine_definition

_f = lib.as_field
_i3 = lib.int3


def _stream():
    return lib.current_stream()


def _call(name, *args):
    lib.check(getattr(lib.load(), name)(*args), name)


# ------------------------------------------------------------------ scheme descriptors
class FluxScheme:
    """What ``get_subroutine_definition('flux_dry' | 'flux_moist')`` returns for b200: a
    descriptor the fused CUDA kernel is specialised on, standing in for
    src/tasmania/isentropic/dynamics/subclasses/minimal_horizontal_fluxes/*.py."""

    def __init__(self, name, extent, order):
        self.name, self.extent, self.order = name, extent, order
        self.code = lib.FLUX_SCHEMES[name]

    def __repr__(self):
        return f"FluxScheme({self.name!r})"


FLUX = {
    "upwind": FluxScheme("upwind", 1, 1),
    "centered": FluxScheme("centered", 1, 2),
    "third_order_upwind": FluxScheme("third_order_upwind", 2, 3),
    "fifth_order_upwind": FluxScheme("fifth_order_upwind", 3, 5),
}


class AdvectionScheme:
    """b200 descriptor of a Burgers advection subroutine
    (src/tasmania/burgers/dynamics/subclasses/advection/*.py)."""

    def __init__(self, name, order):
        self.name, self.order, self.extent = name, order, (order + 1) // 2

    def __repr__(self):
        return f"AdvectionScheme({self.name!r})"


ADVECTION = {
    n: AdvectionScheme(n, o)
    for o, n in enumerate(
        ("first_order", "second_order", "third_order", "fourth_order", "fifth_order", "sixth_order"), 1
    )
}


class SedimentationFluxScheme:
    """b200 descriptor of a sedimentation flux subroutine
    (src/tasmania/physics/microphysics/sedimentation_fluxes/{first,second}_order.py)."""

    def __init__(self, order):
        self.order = int(order)
        self.nb = self.order  # number of upstream levels (first_order.py:L33, second_order.py:L33)

    def __repr__(self):
        return f"SedimentationFluxScheme(order={self.order})"


# class-scoped Kessler stencils for tasmania_b200.plugin: (module, class, stencil, definition name)
KESSLER_CLASS_STENCILS = []


def _scheme_of(obj):
    """A descriptor, or the function object the plugin wraps it in (plugin._descriptor)."""
    return getattr(obj, "tb200_scheme", obj)


def framework_definition(name):
    """The registered b200 definition of a stencil name."""
    from tasmania_b200.framework import get_stencil_definition

    return get_stencil_definition(name)


def _flux_code(externals):
    d = _scheme_of(externals.get("flux_dry"))
    if isinstance(d, FluxScheme):
        return d.code
    if isinstance(d, str):
        return lib.FLUX_SCHEMES[d]
    raise lib.B200Error("externals['flux_dry'] must be a tasmania_b200 FluxScheme descriptor")


# ------------------------------------------------------------------ K12 element-wise
def _ew(op, out, a, b=None, c=None, f=0.0, *, origin, domain):
    _call("tb200_elementwise", lib.ELEMENTWISE_OPS[op], _f(out), _f(a), _f(b), _f(c), float(f),
          _i3(origin), _i3(domain), _stream())


@stencil_definition("copy")
def copy_b200(externals, *, src, dst, origin, domain):
    _ew("copy", dst, src, origin=origin, domain=domain)


@stencil_definition("copychange")
def copychange_b200(externals, *, src, dst, origin, domain):
    _ew("copychange", dst, src, origin=origin, domain=domain)


@stencil_definition("hyperdiffusion")
def hyperdiffusion_b200(externals, *, in_phi, out_phi, alpha, origin, domain):
    """The reference's class-less ``diffusion`` stencil (stencil_definitions/diffusion.py:L31-L55);
    registered under that name by the plugin, under ``hyperdiffusion`` here because the mirrors'
    registry is flat and ``diffusion`` is the dwarfs' class-scoped stencil."""
    _call("tb200_hyperdiffusion", _f(in_phi), _f(out_phi), float(alpha), _i3(origin), _i3(domain),
          _stream())


@stencil_definition("thomas")
def thomas_b200(externals, *, a, b, c, d, out, origin, domain):
    """The global ``thomas`` stencil, framework/subclasses/stencil_definitions/cla.py:L33-L62."""
    _call("tb200_thomas", _f(a), _f(b), _f(c), _f(d), _f(out), _i3(origin), _i3(domain), _stream())


@stencil_definition("abs")
def abs_b200(externals, *, in_field, out_field, origin, domain):
    _ew("abs", out_field, in_field, origin=origin, domain=domain)


@stencil_definition("iabs")
def iabs_b200(externals, *, inout_field, origin, domain):
    _ew("abs", inout_field, inout_field, origin=origin, domain=domain)


@stencil_definition("add")
def add_b200(externals, *, in_a, in_b, out_c, origin, domain):
    _ew("add", out_c, in_a, in_b, origin=origin, domain=domain)


@stencil_definition("iadd")
def iadd_b200(externals, *, inout_a, in_b, origin, domain):
    _ew("add", inout_a, inout_a, in_b, origin=origin, domain=domain)


@stencil_definition("addsub")
def addsub_b200(externals, *, in_a, in_b, in_c, out_d, origin, domain):
    _ew("addsub", out_d, in_a, in_b, in_c, origin=origin, domain=domain)


@stencil_definition("iaddsub")
def iaddsub_b200(externals, *, inout_a, in_b, in_c, origin, domain):
    _ew("iaddsub", inout_a, inout_a, in_b, in_c, origin=origin, domain=domain)


@stencil_definition("clip")
def clip_b200(externals, *, in_field, out_field, origin, domain):
    _ew("clip", out_field, in_field, origin=origin, domain=domain)


@stencil_definition("iclip")
def iclip_b200(externals, *, inout_field, origin, domain):
    _ew("clip", inout_field, inout_field, origin=origin, domain=domain)


@stencil_definition("fma")
def fma_b200(externals, *, in_a, in_b, out_c, f, origin, domain):
    _ew("fma", out_c, in_a, in_b, f=f, origin=origin, domain=domain)


@stencil_definition("mul")
def mul_b200(externals, *, in_a, in_b, out_c, origin, domain):
    _ew("mul", out_c, in_a, in_b, origin=origin, domain=domain)


@stencil_definition("imul")
def imul_b200(externals, *, inout_a, in_b, origin, domain):
    _ew("mul", inout_a, inout_a, in_b, origin=origin, domain=domain)


@stencil_definition("scale")
def scale_b200(externals, *, in_a, out_a, f, origin, domain):
    _ew("scale", out_a, in_a, f=f, origin=origin, domain=domain)


@stencil_definition("iscale")
def iscale_b200(externals, *, inout_a, f, origin, domain):
    _ew("iscale", inout_a, inout_a, f=f, origin=origin, domain=domain)


@stencil_definition("sub")
def sub_b200(externals, *, in_a, in_b, out_c, origin, domain):
    _ew("sub", out_c, in_a, in_b, origin=origin, domain=domain)


@stencil_definition("isub")
def isub_b200(externals, *, inout_a, in_b, origin, domain):
    _ew("sub", inout_a, inout_a, in_b, origin=origin, domain=domain)


FMA_MAX_FIELDS = 8


def fma_fields(outs, ins_a, ins_b, f, *, origin, domain):
    """out_n = a_n + f * b_n over the box for all listed fields, TB200_FMA_MAX_FIELDS per launch:
    one stage of a tendency stepper (tasmania_b200.coupling) instead of one `fma` call per
    field (DataArrayDictOperator.fma, src/tasmania/utils/xarrayx.py:L688-L740)."""
    outs, ins_a, ins_b = list(outs), list(ins_a), list(ins_b)
    if not (len(outs) == len(ins_a) == len(ins_b)):
        raise lib.B200Error("fma_fields: the three field lists differ in length")
    for lo in range(0, len(outs), FMA_MAX_FIELDS):
        keep, arrs = [], []
        for group in (outs, ins_a, ins_b):
            k = [_f(x) for x in group[lo:lo + FMA_MAX_FIELDS]]
            keep.append(k)
            arrs.append((lib.FieldP * len(k))(*[C.pointer(x) for x in k]))
        _call("tb200_fma_fields", len(keep[0]), arrs[0], arrs[1], arrs[2], float(f),
              _i3(origin), _i3(domain), _stream())
        del keep


@stencil_definition("sts_rk2_0")
def sts_rk2_0_b200(externals, *, in_field, in_field_prv, in_tnd, out_field, dt, origin, domain):
    _ew("sts_rk2_0", out_field, in_field, in_field_prv, in_tnd, f=dt, origin=origin, domain=domain)


@stencil_definition("sts_rk3ws_0")
def sts_rk3ws_0_b200(externals, *, in_field, in_field_prv, in_tnd, out_field, dt, origin, domain):
    _ew("sts_rk3ws_0", out_field, in_field, in_field_prv, in_tnd, f=dt, origin=origin, domain=domain)


# ------------------------------------------------------------------ K5 relaxation
@stencil_definition("irelax")
def irelax_b200(externals, *, in_gamma, in_phi_ref, inout_phi, origin, domain):
    _call("tb200_relax", _f(in_gamma), None, _f(in_phi_ref), _f(inout_phi), _i3(origin),
          _i3(domain), _stream())


@stencil_definition("relax")
def relax_b200(externals, *, in_gamma, in_phi, in_phi_ref, out_phi, origin, domain):
    _call("tb200_relax", _f(in_gamma), _f(in_phi), _f(in_phi_ref), _f(out_phi), _i3(origin),
          _i3(domain), _stream())


# ------------------------------------------------------------------ K6 damping
@stencil_definition("damping")
def damping_b200(externals, *, in_phi_now, in_phi_new, in_phi_ref, in_rmat, out_phi, dt, origin,
                 domain):
    _call("tb200_damping", _f(in_phi_now), _f(in_phi_new), _f(in_phi_ref), _f(in_rmat),
          _f(out_phi), float(dt), _i3(origin), _i3(domain), _stream())


# ------------------------------------------------------------------ K4 / K7 diagnostics
@stencil_definition("velocity_x")
def velocity_x_b200(externals, *, in_d, in_du, out_u, origin, domain):
    _call("tb200_velocity", 0, _f(in_d), _f(in_du), _f(out_u), int(bool(externals.get("staggering", True))),
          _i3(origin), _i3(domain), _stream())


@stencil_definition("velocity_y")
def velocity_y_b200(externals, *, in_d, in_dv, out_v, origin, domain):
    _call("tb200_velocity", 1, _f(in_d), _f(in_dv), _f(out_v), int(bool(externals.get("staggering", True))),
          _i3(origin), _i3(domain), _stream())


@stencil_definition("momenta")
def momenta_b200(externals, *, in_d, in_u, in_v, out_du, out_dv, origin, domain):
    _call("tb200_momenta", _f(in_d), _f(in_u), _f(in_v), _f(out_du), _f(out_dv),
          int(bool(externals.get("staggering", True))), _i3(origin), _i3(domain), _stream())


@stencil_definition("density")
def density_b200(externals, *, in_d, in_q, out_dq, origin, domain):
    _call("tb200_density", _f(in_d), _f(in_q), _f(out_dq), int(bool(externals.get("clipping", True))),
          _i3(origin), _i3(domain), _stream())


@stencil_definition("mass_fraction")
def mass_fraction_b200(externals, *, in_d, in_dq, out_q, origin, domain):
    _call("tb200_mass_fraction", _f(in_d), _f(in_dq), _f(out_q),
          int(bool(externals.get("clipping", True))), _i3(origin), _i3(domain), _stream())


# ------------------------------------------------------------------ K8 / K9
def _full_k(field, origin, domain):
    # the reference's numpy diffusion ignores origin[2]/domain[2] and sweeps every level
    # (fourth_order.py:L95-L124 index only i and j)
    return (origin[0], origin[1], 0), (domain[0], domain[1], field.shape[2])


@stencil_definition("diffusion")
def diffusion_b200(externals, *, in_phi, in_gamma, out_phi, dx, dy, ow_out_phi, origin, domain):
    order = externals.get("diffusion_order")
    if order not in (2, 4):
        raise lib.B200Error("externals['diffusion_order'] must be 2 or 4")
    o, d = _full_k(in_phi, origin, domain)
    axis = externals.get("diffusion_axis")  # None: both axes; 0 / 1: the ..._1dx / ..._1dy variants
    if axis is None:
        _call("tb200_diffusion", order, _f(in_phi), _f(in_gamma), _f(out_phi), float(dx), float(dy),
              int(bool(ow_out_phi)), _i3(o), _i3(d), _stream())
    elif axis in (0, 1):
        _call("tb200_diffusion_1d", order, axis, _f(in_phi), _f(in_gamma), _f(out_phi),
              float(dy if axis else dx), int(bool(ow_out_phi)), _i3(o), _i3(d), _stream())
    else:
        raise lib.B200Error("externals['diffusion_axis'] must be None, 0 or 1")


@stencil_definition("smoothing")
def smoothing_b200(externals, *, in_phi, in_gamma, out_phi, origin, domain):
    order = externals.get("smoothing_order")
    if order not in (1, 2, 3):
        raise lib.B200Error("externals['smoothing_order'] must be 1, 2 or 3")
    rim = int(bool(externals.get("rim_copy", False)))
    axis = externals.get("smoothing_axis")
    if axis is None:
        _call("tb200_smoothing", order, _f(in_phi), _f(in_gamma), _f(out_phi), rim,
              _i3(origin), _i3(domain), _stream())
    elif axis in (0, 1):
        _call("tb200_smoothing_1d", order, axis, _f(in_phi), _f(in_gamma), _f(out_phi), rim,
              _i3(origin), _i3(domain), _stream())
    else:
        raise lib.B200Error("externals['smoothing_axis'] must be None, 0 or 1")


# ------------------------------------------------------------------ K1 / K2
def _field_array(fields):
    """array of three ``tb200_field*`` (qv, qc, qr) or NULL"""
    if fields is None or all(f is None for f in fields):
        return None, None
    keep = [_f(x) for x in fields]
    arr = (lib.FieldP * 3)(*[C.pointer(k) if k is not None else None for k in keep])
    return arr, keep


@stencil_definition("step_forward_euler")
def step_forward_euler_b200(
    externals, *, s_now, s_int, s_new, u_int, v_int, su_int=None, sv_int=None, mtg_int=None,
    sqv_now=None, sqv_int=None, sqv_new=None, sqc_now=None, sqc_int=None, sqc_new=None,
    sqr_now=None, sqr_int=None, sqr_new=None, s_tnd=None, qv_tnd=None, qc_tnd=None, qr_tnd=None,
    dt, dx, dy, origin, domain,
):
    moist = bool(externals.get("moist", False))
    # RK3WS-SI passes a zero placeholder when a tendency is absent (rk3ws_si.py:L143); the
    # *_tnd_on switches tell whether it is live
    s_t = s_tnd if externals.get("s_tnd_on", s_tnd is not None) else None
    now = ints = new = tnd = None
    keep = []
    if moist:
        now, k1 = _field_array((sqv_now, sqc_now, sqr_now))
        ints, k2 = _field_array((sqv_int, sqc_int, sqr_int))
        new, k3 = _field_array((sqv_new, sqc_new, sqr_new))
        q = [
            t if externals.get(flag, t is not None) else None
            for t, flag in ((qv_tnd, "qv_tnd_on"), (qc_tnd, "qc_tnd_on"), (qr_tnd, "qr_tnd_on"))
        ]
        tnd, k4 = _field_array(q)
        keep = [k1, k2, k3, k4]
    _call("tb200_step_forward_euler", _flux_code(externals), _f(s_now), _f(s_int), _f(s_new),
          _f(u_int), _f(v_int), _f(s_t), now, ints, new, tnd, float(dt), float(dx), float(dy),
          _i3(origin), _i3(domain), _stream())
    del keep


@stencil_definition("step_forward_euler_momentum")
def step_forward_euler_momentum_b200(
    externals, *, s_now, s_int=None, s_new, u_int, v_int, su_now, su_int, su_new, sv_now, sv_int,
    sv_new, mtg_now, mtg_new, mtg_int=None, su_tnd=None, sv_tnd=None, dt, dx, dy, eps, origin,
    domain,
):
    su_t = su_tnd if externals.get("su_tnd_on", su_tnd is not None) else None
    sv_t = sv_tnd if externals.get("sv_tnd_on", sv_tnd is not None) else None
    _call("tb200_step_forward_euler_momentum", _flux_code(externals), _f(s_now), _f(s_new),
          _f(u_int), _f(v_int), _f(su_now), _f(su_int), _f(su_new), _f(sv_now), _f(sv_int),
          _f(sv_new), _f(mtg_now), _f(mtg_new), _f(su_t), _f(sv_t), float(dt), float(dx),
          float(dy), float(eps), _i3(origin), _i3(domain), _stream())


# ------------------------------------------------------------------ K3
def _constants(externals):
    return lib.Double4(float(externals["pref"]), float(externals["rd"]), float(externals["g"]),
                       float(externals["cp"]))


@stencil_definition("montgomery")
def montgomery_b200(externals, *, in_hs, in_s, inout_mtg, dz, pt, theta_s, origin, domain):
    _call("tb200_montgomery", _f(in_hs), _f(in_s), _f(inout_mtg), float(dz), float(pt),
          float(theta_s), _constants(externals), _i3(origin), _i3(domain), _stream())


@stencil_definition("diagnostic_variables")
def diagnostic_variables_b200(externals, *, in_theta, in_hs, in_s, inout_p, out_exn, inout_mtg,
                              inout_h, dz, pt, origin, domain):
    _call("tb200_diagnostic_variables", _f(in_theta), _f(in_hs), _f(in_s), _f(inout_p),
          _f(out_exn), _f(inout_mtg), _f(inout_h), float(dz), float(pt), _constants(externals),
          _i3(origin), _i3(domain), _stream())


@stencil_definition("height")
def height_b200(externals, *, in_theta, in_hs, in_s, inout_h, dz, pt, origin, domain):
    _call("tb200_height", _f(in_theta), _f(in_hs), _f(in_s), _f(inout_h), float(dz), float(pt),
          _constants(externals), _i3(origin), _i3(domain), _stream())


@stencil_definition("density_and_temperature")
def density_and_temperature_b200(externals, *, in_theta, in_s, in_exn, in_h, out_rho, out_t,
                                 origin, domain):
    _call("tb200_density_and_temperature", _f(in_theta), _f(in_s), _f(in_exn), _f(in_h),
          _f(out_rho), _f(out_t), float(externals["cp"]), _i3(origin), _i3(domain), _stream())


# ------------------------------------------------------------------ K10 Burgers
@stencil_definition("forward_euler")
def burgers_forward_euler_b200(externals, *, in_u, in_v, in_u_tmp, in_v_tmp, out_u, out_v,
                               in_u_tnd=None, in_v_tnd=None, dt, dx, dy, origin, domain):
    adv = _scheme_of(externals.get("advection"))
    order = adv.order if isinstance(adv, AdvectionScheme) else int(adv)
    tu = in_u_tnd if externals.get("tnd_u", in_u_tnd is not None) else None
    tv = in_v_tnd if externals.get("tnd_v", in_v_tnd is not None) else None
    _call("tb200_burgers_forward_euler", order, _f(in_u), _f(in_v), _f(in_u_tmp), _f(in_v_tmp),
          _f(out_u), _f(out_v), _f(tu), _f(tv), float(dt), float(dx), float(dy), _i3(origin),
          _i3(domain), _stream())


# ------------------------------------------------------------------ K11 Kessler
def _kflags(externals, **ow):
    fl = lib.KESSLER_FLAGS
    v = 0
    if externals.get("air_pressure_on_interface_levels", True):
        v |= fl["p_on_interfaces"]
    if externals.get("rain_evaporation", True):
        v |= fl["rain_evaporation"]
    for key, on in ow.items():
        if on:
            v |= fl[key]
    return v


@stencil_definition("kessler")
def kessler_b200(externals, *, in_rho, in_p, in_t, in_exn, in_qc, in_qr, in_qv, out_qc_tnd,
                 out_qr_tnd, out_qv_tnd=None, out_theta_tnd=None, a, k1, k2, ow_out_qc_tnd,
                 ow_out_qr_tnd, ow_out_qv_tnd=True, ow_out_theta_tnd=True, origin, domain):
    flags = _kflags(externals, ow_qc=ow_out_qc_tnd, ow_qr=ow_out_qr_tnd, ow_qv=ow_out_qv_tnd,
                    ow_theta=ow_out_theta_tnd)
    _call("tb200_kessler", _f(in_rho), _f(in_p), _f(in_t), _f(in_exn), _f(in_qc), _f(in_qr),
          _f(in_qv), _f(out_qc_tnd), _f(out_qr_tnd), _f(out_qv_tnd), _f(out_theta_tnd), float(a),
          float(k1), float(k2), float(externals["beta"]), float(externals["lhvw"]), flags,
          _i3(origin), _i3(domain), _stream())


@stencil_definition("saturation_diagnostic")
def saturation_diagnostic_b200(externals, *, in_p, in_t, in_exn, in_qv, in_qc, out_qv, out_qc,
                               out_t, tnd_theta, dt, ow_tnd_theta, origin, domain):
    flags = _kflags(externals, ow_theta=ow_tnd_theta)
    _call("tb200_saturation_diagnostic", _f(in_p), _f(in_t), _f(in_exn), _f(in_qv), _f(in_qc),
          _f(out_qv), _f(out_qc), _f(out_t), _f(tnd_theta), float(dt), float(externals["beta"]),
          float(externals["lhvw"]), float(externals["cp"]), float(externals["rv"]), flags,
          _i3(origin), _i3(domain), _stream())


@stencil_definition("saturation_prognostic")
def saturation_prognostic_b200(externals, *, in_p, in_t, in_exn, in_qv, in_qc, tnd_qv, tnd_qc,
                               tnd_theta, sr, ow_tnd_qv, ow_tnd_qc, ow_tnd_theta, origin, domain):
    flags = _kflags(externals, ow_qv=ow_tnd_qv, ow_qc=ow_tnd_qc, ow_theta=ow_tnd_theta)
    _call("tb200_saturation_prognostic", _f(in_p), _f(in_t), _f(in_exn), _f(in_qv), _f(in_qc),
          _f(tnd_qv), _f(tnd_qc), _f(tnd_theta), float(sr), float(externals["beta"]),
          float(externals["lhvw"]), float(externals["cp"]), float(externals["rv"]), flags,
          _i3(origin), _i3(domain), _stream())


@stencil_definition("fall_velocity")
def fall_velocity_b200(externals, *, in_rho, in_rho_s, in_qr, out_vt, origin, domain):
    _call("tb200_fall_velocity", _f(in_rho), _f(in_rho_s), _f(in_qr), _f(out_vt), _i3(origin),
          _i3(domain), _stream())


@stencil_definition("sedimentation")
def sedimentation_b200(externals, *, in_rho, in_h, in_qr, in_vt, out_tnd_qr, ow_out_tnd_qr, origin,
                       domain):
    sflux = _scheme_of(externals.get("sflux"))
    order = sflux.order if isinstance(sflux, SedimentationFluxScheme) else int(externals["sflux_extent"])
    _call("tb200_sedimentation", order, _f(in_rho), _f(in_h), _f(in_qr), _f(in_vt), _f(out_tnd_qr),
          int(bool(ow_out_tnd_qr)), _i3(origin), _i3(domain), _stream())


@stencil_definition("accumulated_precipitation")
def accumulated_precipitation_b200(externals, *, in_rho, in_qr, in_vt, in_accprec, out_prec,
                                   out_accprec, dt, origin, domain):
    _call("tb200_accumulated_precipitation", _f(in_rho), _f(in_qr), _f(in_vt), _f(in_accprec),
          _f(out_prec), _f(out_accprec), float(dt), float(externals["rhow"]), _i3(origin),
          _i3(domain), _stream())


# ------------------------------------------------------------------ Smagorinsky (8f-3)
@stencil_definition("smagorinsky")
def smagorinsky_b200(externals, *, in_u, in_v, out_u_tnd, out_v_tnd, dx, dy, cs, ow_out_u_tnd,
                     ow_out_v_tnd, origin, domain):
    """Smagorinsky2d's stencil (physics/turbulence.py:L165-L187)."""
    _call("tb200_smagorinsky", None, _f(in_u), _f(in_v), _f(out_u_tnd), _f(out_v_tnd), float(dx),
          float(dy), float(cs), int(bool(ow_out_u_tnd)), int(bool(ow_out_v_tnd)), _i3(origin),
          _i3(domain), _stream())


@stencil_definition("smagorinsky_isentropic")
def smagorinsky_isentropic_b200(externals, *, in_s, in_su, in_sv, out_su_tnd, out_sv_tnd, dx, dy, cs,
                                ow_out_su_tnd, ow_out_sv_tnd, origin, domain):
    """IsentropicSmagorinsky's stencil (isentropic/physics/turbulence.py:L99-L125)."""
    _call("tb200_smagorinsky", _f(in_s), _f(in_su), _f(in_sv), _f(out_su_tnd), _f(out_sv_tnd),
          float(dx), float(dy), float(cs), int(bool(ow_out_su_tnd)), int(bool(ow_out_sv_tnd)),
          _i3(origin), _i3(domain), _stream())


# ------------------------------------------------------------------ Coriolis (8f-3)
@stencil_definition("coriolis")
def coriolis_b200(externals, *, in_su, in_sv, tnd_su, tnd_sv, f, ow_tnd_su, ow_tnd_sv, origin, domain):
    """IsentropicConservativeCoriolis's stencil (isentropic/physics/coriolis.py:L166-L186)."""
    _call("tb200_coriolis", _f(in_su), _f(in_sv), _f(tnd_su), _f(tnd_sv), float(f),
          int(bool(ow_tnd_su)), int(bool(ow_tnd_sv)), _i3(origin), _i3(domain), _stream())


@stencil_definition("coriolis_step")
def coriolis_step_b200(externals, *, in_su, in_sv, base_su, base_sv, out_su, out_sv, f, factor, origin, domain):
    """b200 only: Coriolis forcing fused with the stage update of a tendency stepper
    (``tb200_coriolis_step``): out = base + factor * tendency over the whole storages."""
    _call("tb200_coriolis_step", _f(in_su), _f(in_sv), _f(base_su), _f(base_sv), _f(out_su), _f(out_sv),
          float(f), float(factor), _i3(origin), _i3(domain), _i3(out_su.shape), _stream())


@stencil_definition("smagorinsky_isentropic_step")
def smagorinsky_isentropic_step_b200(externals, *, in_s, in_su, in_sv, base_su, base_sv, out_su, out_sv, dx, dy,
                                     cs, factor, origin, domain):
    """b200 only: IsentropicSmagorinsky fused with the stage update of a tendency stepper
    (``tb200_smagorinsky_step``)."""
    _call("tb200_smagorinsky_step", _f(in_s), _f(in_su), _f(in_sv), _f(base_su), _f(base_sv), _f(out_su),
          _f(out_sv), float(dx), float(dy), float(cs), float(factor), _i3(origin), _i3(domain),
          _i3(out_su.shape), _stream())


# ------------------------------------------------------------------ vertical advection (8f-1)
def _vflux_code(externals):
    d = _scheme_of(externals.get("get_flux_dry", externals.get("flux_dry")))
    if isinstance(d, FluxScheme):
        return d.code
    if isinstance(d, str):
        return lib.FLUX_SCHEMES[d]
    raise lib.B200Error("externals['get_flux_dry'] must be a tasmania_b200 FluxScheme descriptor")


@stencil_definition("vertical_advection")
def vertical_advection_b200(
    externals, *, in_w, in_s, in_su, in_sv, out_s, out_su, out_sv, in_qv=None, in_qc=None,
    in_qr=None, out_qv=None, out_qc=None, out_qr=None, dt=0.0, dz, ow_out_s, ow_out_su, ow_out_sv,
    ow_out_qv=True, ow_out_qc=True, ow_out_qr=True, origin, domain):
    """IsentropicVerticalAdvection's stencil (vertical_advection.py:L271-L386); externals as set
    by the class (L148-L160): get_flux_dry (scheme descriptor), moist, staggering."""
    moist = bool(externals.get("moist", False))
    if moist and any(x is None for x in (in_qv, in_qc, in_qr, out_qv, out_qc, out_qr)):
        raise lib.B200Error("vertical_advection: moist=True needs in_q* and out_q*")
    ows = (ow_out_s, ow_out_su, ow_out_sv, ow_out_qv, ow_out_qc, ow_out_qr)
    flags = sum(1 << n for n, ow in enumerate(ows) if ow)
    q = [(_f(x) if moist else None) for x in (in_qv, in_qc, in_qr, out_qv, out_qc, out_qr)]
    _call("tb200_vertical_advection", _vflux_code(externals), int(bool(externals.get("staggering", False))),
          _f(in_w), _f(in_s), _f(in_su), _f(in_sv), _f(out_s), _f(out_su), _f(out_sv),
          q[0], q[1], q[2], q[3], q[4], q[5], float(dz), flags, _i3(origin), _i3(domain), _stream())


@stencil_definition("vertical_advection_step")
def vertical_advection_step_b200(externals, *, in_w, ins, bases, outs, dz, factor, origin, domain):
    """b200 only: the vertical advection stencil fused with the stage update of a tendency stepper,
    outs[f] = bases[f] + factor * tendency[f](ins) (``tb200_vertical_advection_step``).  Same
    externals as ``vertical_advection``; ins / bases / outs = s, su, sv[, qv, qc, qr]."""
    n = 6 if bool(externals.get("moist", False)) else 3
    if not (len(ins) == len(bases) == len(outs) == n):
        raise lib.B200Error(f"vertical_advection_step: {n} fields expected in ins, bases and outs")
    keep, arrs = [], []
    for group in (ins, bases, outs):
        k = [_f(x) for x in group]
        keep.append(k)
        arrs.append((lib.FieldP * n)(*[C.pointer(x) for x in k]))
    _call("tb200_vertical_advection_step", _vflux_code(externals),
          int(bool(externals.get("staggering", False))), _f(in_w), n, arrs[0], arrs[1], arrs[2],
          float(dz), float(factor), _i3(origin), _i3(domain), _stream())
    del keep


@stencil_definition("implicit_vertical_advection")
def implicit_vertical_advection_b200(
    externals, *, in_w, in_s, in_su, in_sv, out_s, out_su, out_sv, in_qv=None, in_qc=None,
    in_qr=None, out_qv=None, out_qc=None, out_qr=None, gamma, origin, domain):
    """IsentropicImplicitVerticalAdvectionDiagnostic's stencil
    (implicit_vertical_advection.py:L221-L336); externals (L112-L117): moist, staggering."""
    moist = bool(externals.get("moist", False))
    if moist and any(x is None for x in (in_qv, in_qc, in_qr, out_qv, out_qc, out_qr)):
        raise lib.B200Error("implicit_vertical_advection: moist=True needs in_q* and out_q*")
    q = [(_f(x) if moist else None) for x in (in_qv, in_qc, in_qr, out_qv, out_qc, out_qr)]
    _call("tb200_implicit_vertical_advection", int(bool(externals.get("staggering", False))), _f(in_w),
          _f(in_s), _f(in_su), _f(in_sv), _f(out_s), _f(out_su), _f(out_sv), q[0], q[1], q[2], q[3],
          q[4], q[5], float(gamma), 0.0, _i3(origin), _i3(domain), _stream())


@stencil_definition("implicit_vertical_advection_tendency")
def implicit_vertical_advection_tendency_b200(
    externals, *, in_w, in_s, in_su, in_sv, tnd_s, tnd_su, tnd_sv, in_qv=None, in_qc=None,
    in_qr=None, tnd_qv=None, tnd_qc=None, tnd_qr=None, dt, gamma, origin, domain):
    """IsentropicImplicitVerticalAdvectionPrognostic's stencil
    (implicit_vertical_advection.py:L793-L919); externals (L665-L670): moist, vstaggering."""
    moist = bool(externals.get("moist", False))
    if moist and any(x is None for x in (in_qv, in_qc, in_qr, tnd_qv, tnd_qc, tnd_qr)):
        raise lib.B200Error("implicit_vertical_advection: moist=True needs in_q* and tnd_q*")
    if not float(dt) > 0.0:
        raise lib.B200Error("implicit_vertical_advection: dt must be positive")
    q = [(_f(x) if moist else None) for x in (in_qv, in_qc, in_qr, tnd_qv, tnd_qc, tnd_qr)]
    _call("tb200_implicit_vertical_advection", int(bool(externals.get("vstaggering", False))), _f(in_w),
          _f(in_s), _f(in_su), _f(in_sv), _f(tnd_s), _f(tnd_su), _f(tnd_sv), q[0], q[1], q[2], q[3],
          q[4], q[5], float(gamma), float(dt), _i3(origin), _i3(domain), _stream())


_KE = "tasmania.physics.microphysics.kessler"
KESSLER_CLASS_STENCILS += [
    (_KE, "KesslerMicrophysics", "kessler", "kessler_b200"),
    (_KE, "KesslerSaturationAdjustmentDiagnostic", "saturation", "saturation_diagnostic_b200"),
    (_KE, "KesslerSaturationAdjustmentPrognostic", "saturation", "saturation_prognostic_b200"),
    (_KE, "KesslerFallVelocity", "fall_velocity", "fall_velocity_b200"),
    (_KE, "KesslerSedimentation", "sedimentation", "sedimentation_b200"),
    ("tasmania.physics.microphysics.utils", "Precipitation", "accumulated_precipitation",
     "accumulated_precipitation_b200"),
]
SEDIMENTATION_FLUX = {"first_order_upwind": SedimentationFluxScheme(1),
                      "second_order_upwind": SedimentationFluxScheme(2)}


# subroutine descriptors: must *exist* for the backend (stencil.py:L379-L392)
for _name, _scheme in FLUX.items():
    subroutine_definition(f"flux_dry:{_name}")(_scheme)
    subroutine_definition(f"flux_moist:{_name}")(_scheme)
for _name, _scheme in ADVECTION.items():
    subroutine_definition(f"advection:{_name}")(_scheme)
for _name, _scheme in SEDIMENTATION_FLUX.items():
    subroutine_definition(f"flux:{_name}")(_scheme)
subroutine_definition("set_output")("set_output")
subroutine_definition("smagorinsky_core")("smagorinsky_core")  # fused into tb200_smagorinsky
for _name in ("thomas", "setup_thomas", "setup_thomas_bc"):  # fused into tb200_implicit_vertical_advection
    subroutine_definition(_name)(_name)
