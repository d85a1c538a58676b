// kessler.cu -- K11: Kessler warm-rain microphysics, saturation adjustment, raindrop fall
// velocity, sedimentation and accumulated precipitation.
//
// Reference (numpy definitions):
//   src/tasmania/physics/microphysics/kessler.py:L307-L376     kessler
//   src/tasmania/physics/microphysics/kessler.py:L661-L714     saturation (diagnostic adjustment)
//   src/tasmania/physics/microphysics/kessler.py:L981-L1032    saturation (prognostic)
//   src/tasmania/physics/microphysics/kessler.py:L1183-L1203   fall_velocity
//   src/tasmania/physics/microphysics/kessler.py:L1339-L1370   sedimentation
//   src/tasmania/physics/microphysics/sedimentation_fluxes/{first,second}_order.py:L36-L59
//   src/tasmania/physics/microphysics/utils.py:L283-L305       accumulated_precipitation
//
// All are point-wise in (i, j) with at most k+1 / k-1 / k-2 neighbours: one thread per point,
// i along the warp, pure streams (88 / 64 / 24 / 40 B per point).  Operation order is the
// reference's; exp is CUDA libm (<= 2 ulp, glibc/numpy <= 1 ulp), the powers of positive arguments
// go through the kernels' own pow_pos (stencil_math.cuh: <= 1.7e-16 relative, a third of the
// instructions of CUDA's generic pow; anything outside its domain falls back to that),
// so these stencils are held to 1e-13 relative instead of bit-exactness.  x ** 0.5 and x ** 2.0
// are sqrt and x * x in numpy (scalar-power fast paths) and here.
#include "stencil_math.cuh"

using namespace tb200;

namespace {

__device__ __forceinline__ void set_output(double &lhs, double rhs, bool overwrite) {
  lhs = overwrite ? rhs : lhs + rhs;  // generics.py:L38-L40
}

// Tetens' formula and the saturation mixing ratio, kessler.py:L345-L348
__device__ __forceinline__ double saturation_mixing_ratio(double t, double p, double beta) {
  const double ps = 610.78 * exp(17.27 * (t - 273.16) / (t - 35.86));
  return beta * ps / p;
}

__device__ __forceinline__ void main_level(const View &p, const View &exn, int i, int j, int k,
                                           bool on_interfaces, double &pm, double &em) {
  if (on_interfaces) {
    pm = 0.5 * (p(i, j, k) + p(i, j, k + 1));
    em = 0.5 * (exn(i, j, k) + exn(i, j, k + 1));
  } else {
    pm = p(i, j, k);
    em = exn(i, j, k);
  }
}

}  // namespace

extern "C" int tb200_kessler(const tb200_field *in_rho, const tb200_field *in_p,
                             const tb200_field *in_t, const tb200_field *in_exn,
                             const tb200_field *in_qc, const tb200_field *in_qr,
                             const tb200_field *in_qv, tb200_field *out_qc_tnd,
                             tb200_field *out_qr_tnd, tb200_field *out_qv_tnd,
                             tb200_field *out_theta_tnd, double a, double k1, double k2,
                             double beta, double lhvw, uint32_t flags, const int32_t origin[3],
                             const int32_t domain[3], void *stream) {
  View rho = view(in_rho), p = view(in_p), t = view(in_t), exn = view(in_exn), qc = view(in_qc),
       qr = view(in_qr), qv = view(in_qv), tqc = view(out_qc_tnd), tqr = view(out_qr_tnd),
       tqv = view(out_qv_tnd), tth = view(out_theta_tnd);
  const bool apoil = flags & TB200_KESSLER_P_ON_INTERFACES, evap = flags & TB200_KESSLER_RAIN_EVAPORATION;
  const bool ow_qc = flags & TB200_KESSLER_OW_QC, ow_qr = flags & TB200_KESSLER_OW_QR,
             ow_qv = flags & TB200_KESSLER_OW_QV, ow_th = flags & TB200_KESSLER_OW_THETA;
  const int hk = apoil ? 1 : 0;
  TB200_REQUIRE(box_inside(rho, origin, domain) && box_inside(t, origin, domain) &&
                    box_inside(qc, origin, domain) && box_inside(qr, origin, domain) &&
                    box_inside(qv, origin, domain) && box_inside(tqc, origin, domain) &&
                    box_inside(tqr, origin, domain),
                "kessler: box outside a storage");
  TB200_REQUIRE(box_inside(p, origin, domain, 0, 0, 0, 0, 0, hk) &&
                    box_inside(exn, origin, domain, 0, 0, 0, 0, 0, hk),
                "kessler: pressure / Exner box outside the storage");
  TB200_REQUIRE(!evap || (box_inside(tqv, origin, domain) && box_inside(tth, origin, domain)),
                "kessler: rain evaporation needs out_qv_tnd and out_theta_tnd");
  const int i0 = origin[0], j0 = origin[1], k0 = origin[2];
  return launch_box("kessler", domain, static_cast<cudaStream_t>(stream),
                    [=] __device__(int i, int j, int k) {
                      i += i0; j += j0; k += k0;
                      double pm, em;
                      main_level(p, exn, i, j, k, apoil, pm, em);
                      const double c = qc(i, j, k), r = qr(i, j, k);
                      const double ar = k1 * (c > a ? c - a : 0.0);
                      const double cr = k2 * c * (r > 0.0 ? pow_pos(r, 0.875) : 0.0);
                      if (!evap) {
                        set_output(tqc(i, j, k), -(ar + cr), ow_qc);
                        set_output(tqr(i, j, k), ar + cr, ow_qr);
                        return;
                      }
                      const double qvs = saturation_mixing_ratio(t(i, j, k), pm, beta);
                      const double er =
                          r > 0.0 ? 0.0484794 * (qvs - qv(i, j, k)) * pow_pos(rho(i, j, k) * r, 13.0 / 20.0)
                                  : 0.0;
                      set_output(tqv(i, j, k), er, ow_qv);
                      set_output(tqc(i, j, k), -(ar + cr), ow_qc);
                      set_output(tqr(i, j, k), ar + cr - er, ow_qr);
                      set_output(tth(i, j, k), -lhvw / em * er, ow_th);
                    });
}

// MODE 0: diagnostic adjustment (outputs qv, qc, t and the theta tendency);
// MODE 1: prognostic (tendencies of qv, qc, theta scaled by the saturation rate)
template <int MODE>
static int run_saturation(View p, View t, View exn, View qv, View qc, View o0, View o1, View o2,
                          View tth, double x, double beta, double lhvw, double cp, double rv,
                          uint32_t flags, const int32_t o[3], const int32_t d[3], cudaStream_t st) {
  const bool apoil = flags & TB200_KESSLER_P_ON_INTERFACES;
  const bool ow_qv = flags & TB200_KESSLER_OW_QV, ow_qc = flags & TB200_KESSLER_OW_QC,
             ow_th = flags & TB200_KESSLER_OW_THETA;
  const int i0 = o[0], j0 = o[1], k0 = o[2];
  const double lhvw2 = lhvw * lhvw;  // lhvw ** 2.0
  return launch_box("saturation", d, st, [=] __device__(int i, int j, int k) {
    i += i0; j += j0; k += k0;
    double pm, em;
    main_level(p, exn, i, j, k, apoil, pm, em);
    const double tk = t(i, j, k), v = qv(i, j, k), c = qc(i, j, k);
    const double qvs = saturation_mixing_ratio(tk, pm, beta);
    const double sat = (qvs - v) / (1.0 + qvs * lhvw2 / (cp * rv * (tk * tk)));
    const double dq = sat <= c ? sat : c;
    if (MODE == 0) {
      o0(i, j, k) = v + dq;
      o1(i, j, k) = c - dq;
      o2(i, j, k) = tk - dq * lhvw / cp;
      set_output(tth(i, j, k), (lhvw / em) * (-dq / x), ow_th);  // x = dt
    } else {
      set_output(o0(i, j, k), x * dq, ow_qv);  // x = saturation rate
      set_output(o1(i, j, k), -x * dq, ow_qc);
      set_output(tth(i, j, k), -x * (lhvw / em) * dq, ow_th);
    }
  });
}

extern "C" int tb200_saturation_diagnostic(
    const tb200_field *in_p, const tb200_field *in_t, const tb200_field *in_exn,
    const tb200_field *in_qv, const tb200_field *in_qc, tb200_field *out_qv, tb200_field *out_qc,
    tb200_field *out_t, tb200_field *tnd_theta, double dt, double beta, double lhvw, double cp,
    double rv, uint32_t flags, const int32_t origin[3], const int32_t domain[3], void *stream) {
  View p = view(in_p), t = view(in_t), exn = view(in_exn), qv = view(in_qv), qc = view(in_qc),
       oqv = view(out_qv), oqc = view(out_qc), ot = view(out_t), tth = view(tnd_theta);
  const int hk = (flags & TB200_KESSLER_P_ON_INTERFACES) ? 1 : 0;
  TB200_REQUIRE(box_inside(t, origin, domain) && box_inside(qv, origin, domain) &&
                    box_inside(qc, origin, domain) && box_inside(oqv, origin, domain) &&
                    box_inside(oqc, origin, domain) && box_inside(ot, origin, domain) &&
                    box_inside(tth, origin, domain) &&
                    box_inside(p, origin, domain, 0, 0, 0, 0, 0, hk) &&
                    box_inside(exn, origin, domain, 0, 0, 0, 0, 0, hk),
                "saturation (diagnostic): box outside a storage");
  return run_saturation<0>(p, t, exn, qv, qc, oqv, oqc, ot, tth, dt, beta, lhvw, cp, rv, flags,
                           origin, domain, static_cast<cudaStream_t>(stream));
}

extern "C" int tb200_saturation_prognostic(
    const tb200_field *in_p, const tb200_field *in_t, const tb200_field *in_exn,
    const tb200_field *in_qv, const tb200_field *in_qc, tb200_field *tnd_qv, tb200_field *tnd_qc,
    tb200_field *tnd_theta, double sr, double beta, double lhvw, double cp, double rv,
    uint32_t flags, const int32_t origin[3], const int32_t domain[3], void *stream) {
  View p = view(in_p), t = view(in_t), exn = view(in_exn), qv = view(in_qv), qc = view(in_qc),
       tqv = view(tnd_qv), tqc = view(tnd_qc), tth = view(tnd_theta);
  const int hk = (flags & TB200_KESSLER_P_ON_INTERFACES) ? 1 : 0;
  TB200_REQUIRE(box_inside(t, origin, domain) && box_inside(qv, origin, domain) &&
                    box_inside(qc, origin, domain) && box_inside(tqv, origin, domain) &&
                    box_inside(tqc, origin, domain) && box_inside(tth, origin, domain) &&
                    box_inside(p, origin, domain, 0, 0, 0, 0, 0, hk) &&
                    box_inside(exn, origin, domain, 0, 0, 0, 0, 0, hk),
                "saturation (prognostic): box outside a storage");
  return run_saturation<1>(p, t, exn, qv, qc, tqv, tqc, View{}, tth, sr, beta, lhvw, cp, rv, flags,
                           origin, domain, static_cast<cudaStream_t>(stream));
}

extern "C" int tb200_fall_velocity(const tb200_field *in_rho, const tb200_field *in_rho_s,
                                   const tb200_field *in_qr, tb200_field *out_vt,
                                   const int32_t origin[3], const int32_t domain[3], void *stream) {
  View rho = view(in_rho), rho_s = view(in_rho_s), qr = view(in_qr), vt = view(out_vt);
  TB200_REQUIRE(box_inside(rho, origin, domain) && box_inside(rho_s, origin, domain) &&
                    box_inside(qr, origin, domain) && box_inside(vt, origin, domain),
                "fall_velocity: box outside a storage");
  const int i0 = origin[0], j0 = origin[1], k0 = origin[2];
  return launch_box("fall_velocity", domain, static_cast<cudaStream_t>(stream),
                    [=] __device__(int i, int j, int k) {
                      i += i0; j += j0; k += k0;
                      const double r = qr(i, j, k), d = rho(i, j, k);
                      const double w = 1.0e-3 * d * (r > 0.0 ? r : 0.0);
                      vt(i, j, k) = 36.34 * (w > 0.0 ? pow_pos(w, 0.1346) : 0.0) * sqrt(rho_s(i, j, k) / d);
                    });
}

template <int ORDER>
static int run_sedimentation(View rho, View h, View qr, View vt, View tnd, bool ow,
                             const int32_t o[3], const int32_t d[3], cudaStream_t st) {
  const int i0 = o[0], j0 = o[1], kb = o[2];
  return launch_box("sedimentation", d, st, [=] __device__(int i, int j, int k) {
    i += i0; j += j0; k += kb;
    if (k < kb + ORDER) {  // kessler.py:L1360-L1361
      if (ow) tnd(i, j, k) = 0.0;
      return;
    }
    // heights of the main levels, kessler.py:L1353
    const double h0 = 0.5 * (h(i, j, k) + h(i, j, k + 1));
    const double h1 = 0.5 * (h(i, j, k - 1) + h(i, j, k));
    const double f0 = rho(i, j, k), f1 = rho(i, j, k - 1);
    double dfdz;
    if (ORDER == 1) {  // first_order.py:L38-L42
      dfdz = (f1 * qr(i, j, k - 1) * vt(i, j, k - 1) - f0 * qr(i, j, k) * vt(i, j, k)) / (h1 - h0);
    } else {  // second_order.py:L42-L59
      const double h2 = 0.5 * (h(i, j, k - 2) + h(i, j, k - 1));
      const double a = (2.0 * h0 - h1 - h2) / ((h1 - h0) * (h2 - h0));
      const double b = (h2 - h0) / ((h1 - h0) * (h2 - h1));
      const double c = (h0 - h1) / ((h2 - h0) * (h2 - h1));
      dfdz = a * f0 * qr(i, j, k) * vt(i, j, k) + b * f1 * qr(i, j, k - 1) * vt(i, j, k - 1) +
             c * rho(i, j, k - 2) * qr(i, j, k - 2) * vt(i, j, k - 2);
    }
    set_output(tnd(i, j, k), qdiv(dfdz, f0), ow);  // (same bits as `/`; zero wherever there is no rain)
  });
}

extern "C" int tb200_sedimentation(int order, const tb200_field *in_rho, const tb200_field *in_h,
                                   const tb200_field *in_qr, const tb200_field *in_vt,
                                   tb200_field *out_tnd_qr, int ow_out_tnd_qr,
                                   const int32_t origin[3], const int32_t domain[3], void *stream) {
  View rho = view(in_rho), h = view(in_h), qr = view(in_qr), vt = view(in_vt), tnd = view(out_tnd_qr);
  TB200_REQUIRE(order == 1 || order == 2, "sedimentation: order must be 1 or 2, got %d", order);
  TB200_REQUIRE(box_inside(rho, origin, domain) && box_inside(qr, origin, domain) &&
                    box_inside(vt, origin, domain) && box_inside(tnd, origin, domain) &&
                    box_inside(h, origin, domain, 0, 0, 0, 0, 0, 1),
                "sedimentation: box outside a storage");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return order == 1 ? run_sedimentation<1>(rho, h, qr, vt, tnd, ow_out_tnd_qr != 0, origin, domain, st)
                    : run_sedimentation<2>(rho, h, qr, vt, tnd, ow_out_tnd_qr != 0, origin, domain, st);
}

extern "C" int tb200_accumulated_precipitation(
    const tb200_field *in_rho, const tb200_field *in_qr, const tb200_field *in_vt,
    const tb200_field *in_accprec, tb200_field *out_prec, tb200_field *out_accprec, double dt,
    double rhow, const int32_t origin[3], const int32_t domain[3], void *stream) {
  View rho = view(in_rho), qr = view(in_qr), vt = view(in_vt), acc_in = view(in_accprec),
       prec = view(out_prec), acc = view(out_accprec);
  TB200_REQUIRE(box_inside(rho, origin, domain) && box_inside(qr, origin, domain) &&
                    box_inside(vt, origin, domain) && box_inside(acc_in, origin, domain) &&
                    box_inside(prec, origin, domain) && box_inside(acc, origin, domain),
                "accumulated_precipitation: box outside a storage");
  const int i0 = origin[0], j0 = origin[1], k0 = origin[2];
  return launch_box("accumulated_precipitation", domain, static_cast<cudaStream_t>(stream),
                    [=] __device__(int i, int j, int k) {
                      i += i0; j += j0; k += k0;
                      const double pr = 3.6e6 * rho(i, j, k) * qr(i, j, k) * vt(i, j, k) / rhow;
                      prec(i, j, k) = pr;
                      acc(i, j, k) = acc_in(i, j, k) + dt * pr / 3.6e3;
                    });
}
