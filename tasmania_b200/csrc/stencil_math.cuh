// stencil_math.cuh -- point formulas shared by the stand-alone and the fused kernels.
// Operation order follows the reference's numpy definitions exactly (file:line cited at
// each formula); together with -fmad=false this makes the kernels bit-identical to numpy
// apart from libm calls.
#pragma once
#include "common.cuh"

namespace tb200 {

// ------------------------------------------------------------------ horizontal fluxes
// Face flux F(f) = w[f] * Phi(phi[f-e .. f+e-1]); `ph` points at phi[f], `st` is the stride
// along the differenced axis.  SURVEY.md Appendix A "Flux-array offset".
//
// The high-order schemes scale the advecting velocity once, wq = w / 12 or w / 60; since
// |w| / c == |w / c| bit for bit, both occurrences of the division in the reference formula
// are served by one constant division, and wq is shared by every field advected through the
// same face.
struct FluxConst {
  CDiv c12, c60, dx, dy;
};
inline FluxConst make_flux_const(double dx, double dy) {
  return FluxConst{make_cdiv(12.0), make_cdiv(60.0), make_cdiv(dx), make_cdiv(dy)};
}

template <int SCHEME>
struct Flux;

template <>
struct Flux<TB200_FLUX_UPWIND> {  // horizontal_fluxes/upwind.py:L32-L39
  static constexpr int extent = 1;
  __device__ __forceinline__ static double prep(double w, const FluxConst &) { return w; }
  __device__ __forceinline__ static double face(double w, const double *ph, long long st) {
    return w * (w > 0.0 ? __ldg(ph - st) : __ldg(ph));
  }
  // v[0 .. 2e-1] = phi[f-e .. f+e-1]
  __device__ __forceinline__ static double eval_v(double w, const double *v) {
    return w * (w > 0.0 ? v[0] : v[1]);
  }
};

template <>
struct Flux<TB200_FLUX_CENTERED> {  // horizontal_fluxes/centered.py:L173-L202
  static constexpr int extent = 1;
  __device__ __forceinline__ static double prep(double w, const FluxConst &) { return w * 0.5; }
  __device__ __forceinline__ static double face(double wq, const double *ph, long long st) {
    return wq * (__ldg(ph - st) + __ldg(ph));
  }
  __device__ __forceinline__ static double eval_v(double wq, const double *v) {
    return wq * (v[0] + v[1]);
  }
};

template <>
struct Flux<TB200_FLUX_THIRD_ORDER_UPWIND> {  // horizontal_fluxes/third_order_upwind.py:L32-L55
  static constexpr int extent = 2;
  __device__ __forceinline__ static double prep(double w, const FluxConst &c) { return w / c.c12; }
  __device__ __forceinline__ static double eval(double wq, double m2, double m1, double p0,
                                                double p1) {
    const double flux4 = wq * (7.0 * (p0 + m1) - (p1 + m2));
    return flux4 - fabs(wq) * (3.0 * (p0 - m1) - (p1 - m2));
  }
  __device__ __forceinline__ static double face(double wq, const double *ph, long long st) {
    return eval(wq, __ldg(ph - 2 * st), __ldg(ph - st), __ldg(ph), __ldg(ph + st));
  }
  __device__ __forceinline__ static double eval_v(double wq, const double *v) {
    return eval(wq, v[0], v[1], v[2], v[3]);
  }
};

template <>
struct Flux<TB200_FLUX_FIFTH_ORDER_UPWIND> {  // horizontal_fluxes/fifth_order_upwind.py:L32-L75
  static constexpr int extent = 3;
  __device__ __forceinline__ static double prep(double w, const FluxConst &c) { return w / c.c60; }
  __device__ __forceinline__ static double eval(double wq, double m3, double m2, double m1,
                                                double p0, double p1, double p2) {
    const double flux6 = wq * (37.0 * (p0 + m1) - 8.0 * (p1 + m2) + (p2 + m3));
    return flux6 - fabs(wq) * (10.0 * (p0 - m1) - 5.0 * (p1 - m2) + (p2 - m3));
  }
  __device__ __forceinline__ static double face(double wq, const double *ph, long long st) {
    return eval(wq, __ldg(ph - 3 * st), __ldg(ph - 2 * st), __ldg(ph - st), __ldg(ph),
                __ldg(ph + st), __ldg(ph + 2 * st));
  }
  __device__ __forceinline__ static double eval_v(double wq, const double *v) {
    return eval(wq, v[0], v[1], v[2], v[3], v[4], v[5]);
  }
};

// the four (pre-scaled) face velocities around mass point (i, j, k)
struct FaceVel {
  double xm, xp, ym, yp;
};
template <int SCHEME>
__device__ __forceinline__ FaceVel face_velocities(const View &u, const View &v, int i, int j,
                                                   int k, const FluxConst &c) {
  using F = Flux<SCHEME>;
  return FaceVel{F::prep(u.ld(i, j, k), c), F::prep(u.ld(i + 1, j, k), c),
                 F::prep(v.ld(i, j, k), c), F::prep(v.ld(i, j + 1, k), c)};
}

// (Fx[i+1/2] - Fx[i-1/2]) / dx + (Fy[j+1/2] - Fy[j-1/2]) / dy at mass point (i, j, k),
// prognostics/utils.py:L96-L99.
template <int SCHEME>
__device__ __forceinline__ double flux_divergence(const FaceVel &w, const View &phi, int i,
                                                  int j, int k, const FluxConst &c) {
  using F = Flux<SCHEME>;
  const double *pc = phi.p + (i * phi.s0 + j * phi.s1 + k * phi.s2);
  const double fxm = F::face(w.xm, pc, phi.s0);
  const double fxp = F::face(w.xp, pc + phi.s0, phi.s0);
  const double fym = F::face(w.ym, pc, phi.s1);
  const double fyp = F::face(w.yp, pc + phi.s1, phi.s1);
  return (fxp - fxm) / c.dx + (fyp - fym) / c.dy;
}

// ------------------------------------------------------------------ Exner function
// exn = cp (p / pref)^kappa (isentropic/dynamics/diagnostics.py:L345).  For the positive
// arguments met here x^kappa = exp2(kappa * log2(x)); CUDA's log2 and exp2 are 1-ulp
// functions, which bounds the result's error by ~2 ulp -- the bound CUDA documents for its
// own pow -- at less than half the instructions (pow spends most of its time on sign /
// integer-exponent / infinity special cases).  The reference's glibc pow is < 1 ulp, so either
// way the last bit may differ; parity tests hold K3 to 1e-13 and the 100-step run to 1e-12.
__device__ __forceinline__ double pow_pos(double x, double kappa) {
  return exp2(kappa * log2(x));
}

// ------------------------------------------------------------------ relaxation / damping
// algorithms.py:L32-L43
__device__ __forceinline__ double relax_point(double g, double phi, double ref) {
  return g == 0.0 ? phi : (g == 1.0 ? ref : phi - g * (phi - ref));
}

// rayleigh.py:L104-L109: new - (dt * R) * (now - ref)
__device__ __forceinline__ double damp_point(double now, double nw, double ref, double r,
                                             double dt) {
  return nw - dt * r * (now - ref);
}

// ------------------------------------------------------------------ Burgers advection
// burgers/dynamics/subclasses/advection/{first..sixth}_order.py.  `aq` is the advecting
// velocity at the point already divided by the scheme's denominator (2 d, 12 d or 60 d --
// `Advection<ORDER>::denominator(d)`); |a| / den == |a / den| bit for bit.  `q` points at
// the advected field at the point, `st` is the stride along the differenced axis.
template <int ORDER>
struct Advection;

template <>
struct Advection<1> {  // first_order.py:L39-L58
  static constexpr int extent = 1;
  static double denominator(double d) { return 2.0 * d; }
  __device__ __forceinline__ static double term(double aq, const double *q, long long st) {
    const double qm = __ldg(q - st), q0 = __ldg(q), qp = __ldg(q + st);
    return aq * (qp - qm) - fabs(aq) * (qp - 2.0 * q0 + qm);
  }
};
template <>
struct Advection<2> {  // second_order.py:L37-L45
  static constexpr int extent = 1;
  static double denominator(double d) { return 2.0 * d; }
  __device__ __forceinline__ static double term(double aq, const double *q, long long st) {
    return aq * (__ldg(q + st) - __ldg(q - st));
  }
};
template <>
struct Advection<3> {  // third_order.py:L39-L65
  static constexpr int extent = 2;
  static double denominator(double d) { return 12.0 * d; }
  __device__ __forceinline__ static double term(double aq, const double *q, long long st) {
    const double m2 = __ldg(q - 2 * st), m1 = __ldg(q - st), q0 = __ldg(q);
    const double p1 = __ldg(q + st), p2 = __ldg(q + 2 * st);
    return aq * (8.0 * (p1 - m1) - (p2 - m2)) +
           fabs(aq) * (p2 + m2 - 4.0 * (p1 + m1) + 6.0 * q0);
  }
};
template <>
struct Advection<4> {  // fourth_order.py:L37-L61
  static constexpr int extent = 2;
  static double denominator(double d) { return 12.0 * d; }
  __device__ __forceinline__ static double term(double aq, const double *q, long long st) {
    const double m2 = __ldg(q - 2 * st), m1 = __ldg(q - st);
    const double p1 = __ldg(q + st), p2 = __ldg(q + 2 * st);
    return aq * (8.0 * (p1 - m1) - (p2 - m2));
  }
};
template <>
struct Advection<5> {  // fifth_order.py:L39-L86
  static constexpr int extent = 3;
  static double denominator(double d) { return 60.0 * d; }
  __device__ __forceinline__ static double term(double aq, const double *q, long long st) {
    const double m3 = __ldg(q - 3 * st), m2 = __ldg(q - 2 * st), m1 = __ldg(q - st);
    const double q0 = __ldg(q);
    const double p1 = __ldg(q + st), p2 = __ldg(q + 2 * st), p3 = __ldg(q + 3 * st);
    return aq * (45.0 * (p1 - m1) - 9.0 * (p2 - m2) + (p3 - m3)) -
           fabs(aq) * ((p3 + m3) - 6.0 * (p2 + m2) + 15.0 * (p1 + m1) - 20.0 * q0);
  }
};
template <>
struct Advection<6> {  // sixth_order.py:L37-L77
  static constexpr int extent = 3;
  static double denominator(double d) { return 60.0 * d; }
  __device__ __forceinline__ static double term(double aq, const double *q, long long st) {
    const double m3 = __ldg(q - 3 * st), m2 = __ldg(q - 2 * st), m1 = __ldg(q - st);
    const double p1 = __ldg(q + st), p2 = __ldg(q + 2 * st), p3 = __ldg(q + 3 * st);
    return aq * (45.0 * (p1 - m1) - 9.0 * (p2 - m2) + (p3 - m3));
  }
};

}  // namespace tb200
