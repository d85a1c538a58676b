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
// exn = cp (p / pref)^kappa (isentropic/dynamics/diagnostics.py:L345, L433-L438).  The column
// scans spend most of their instructions in this power, so it gets a domain-specific
// evaluation x^kappa = 2^(kappa log2 x) for positive, normal, finite x (anything else takes
// CUDA's generic pow):
//   * log2: x = 2^e m with m in [sqrt(1/2), sqrt(2)), s = (m - 1) / (m + 1) through one
//     MUFU.RCP64H seed + one cubic Newton step + one residual correction,
//     log m = 2 s + s z P(z), z = s^2, nine Taylor terms (z < 0.0295), then
//     log2 x = e + log2(e) log m kept as an unevaluated sum hi + lo;
//   * y = kappa (hi + lo) with the rounding error of the product recovered by one FMA, so the
//     argument of 2^y carries no rounding of its own (the dominant error of the plain
//     exp2(kappa * log2(x)) formulation);
//   * 2^y = 2^n 2^r, n = rint(y), |r| <= 1/2, degree-13 Taylor polynomial, exponent patched in.
// About 60 instructions (45 fp64) against ~110 for exp2(kappa * log2(x)) and ~250 for pow.
// Measured against 60-digit arithmetic (tests/test_host_setup.py restates the sequence
// operation by operation): relative error <= 1.7e-16, next to 1.3e-16 for glibc's pow that the
// reference calls, so K3 stays within 1e-13 and the 100-step run within 1e-12 of the reference.  Every FMA below is explicit:
// the file is compiled with -fmad=false.
static __device__ __noinline__ double pow_generic(double x, double kappa) { return pow(x, kappa); }

// coefficients as constant-bank operands of the DFMAs (immediates would cost two UMOVs each)
static __constant__ double kPowLog[9] = {  // 2 / (2 k + 3)
    0x1.5555555555555p-1, 0x1.999999999999ap-2, 0x1.2492492492492p-2, 0x1.c71c71c71c71cp-3,
    0x1.745d1745d1746p-3, 0x1.3b13b13b13b14p-3, 0x1.1111111111111p-3, 0x1.e1e1e1e1e1e1ep-4,
    0x1.af286bca1af28p-4};
static __constant__ double kPowExp[14] = {  // ln(2)^k / k!
    1.0,                  0x1.62e42fefa39efp-1,  0x1.ebfbdff82c58fp-3,  0x1.c6b08d704a0c0p-5,
    0x1.3b2ab6fba4e77p-7, 0x1.5d87fe78a6731p-10, 0x1.430912f86c787p-13, 0x1.ffcbfc588b0c7p-17,
    0x1.62c0223a5c824p-20, 0x1.b5253d395e7c4p-24, 0x1.e4cf5158b8ecap-28, 0x1.e8cac7351bb25p-32,
    0x1.c3bd650fc2986p-36, 0x1.816193166d0f9p-40};
static __constant__ double kPowMisc[2] = {0x1.71547652b82fep+0, 6755399441055744.0};  // log2(e), 1.5 * 2^52

// N independent powers, every step written across the N arguments so that the N dependent
// fp64 chains are interleaved by the scheduler (a DFMA chain alone leaves the pipe idle most of
// the time; measured on B200: stall reason "wait" 46 % with one chain per warp).  The fast path
// is branch-free; arguments outside it (<= 0, subnormal, inf, nan, result outside the normal
// range) are patched afterwards by the generic pow.
template <int N>
__device__ __forceinline__ void pow_pos_n(const double (&x)[N], double kappa, double (&out)[N]) {
  double m[N], ef[N], s[N], z[N], p[N], y[N], q[N], w[N];
  int n2[N];
  bool bad = false;
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const int hi = __double2hiint(x[n]);
    bad |= (unsigned)(hi - 0x00100000) >= 0x7fe00000u;
    const int e = (hi - 0x3fe6a09f) >> 20;  // floor(log2(x / sqrt(1/2)))
    m[n] = __hiloint2double(hi - (e << 20), __double2loint(x[n]));
    ef[n] = (double)e;
  }
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const double f = m[n] - 1.0, d = m[n] + 1.0;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double t = fma(-d, r, 1.0);
    t = fma(t, t, t);
    r = fma(r, t, r);
    const double s0 = f * r;
    s[n] = fma(fma(-d, s0, f), r, s0);
    z[n] = s[n] * s[n];
    p[n] = kPowLog[8];
  }
#pragma unroll
  for (int c = 7; c >= 0; --c)
#pragma unroll
    for (int n = 0; n < N; ++n) p[n] = fma(p[n], z[n], kPowLog[c]);
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const double lm = fma(s[n] * z[n], p[n], 2.0 * s[n]);  // log m
    const double l2 = fma(lm, kPowMisc[0], ef[n]);
    const double l2_lo = fma(lm, kPowMisc[0], ef[n] - l2);
    y[n] = kappa * l2;
    const double y_lo = fma(kappa, l2_lo, fma(kappa, l2, -y[n]));
    bad |= !(fabs(y[n]) < 1000.0);
    const double ts = y[n] + kPowMisc[1];  // rint by addition
    n2[n] = __double2loint(ts);
    q[n] = (y[n] - (ts - kPowMisc[1])) + y_lo;
    w[n] = kPowExp[13];
  }
#pragma unroll
  for (int c = 12; c >= 0; --c)
#pragma unroll
    for (int n = 0; n < N; ++n) w[n] = fma(w[n], q[n], kPowExp[c]);
#pragma unroll
  for (int n = 0; n < N; ++n)
    out[n] = __hiloint2double(__double2hiint(w[n]) + (n2[n] << 20), __double2loint(w[n]));
  if (bad) {
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const int hi = __double2hiint(x[n]);
      if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u || !(fabs(y[n]) < 1000.0))
        out[n] = pow_generic(x[n], kappa);
    }
  }
}

__device__ __forceinline__ double pow_pos(double x, double kappa) {
  const double xs[1] = {x};
  double r[1];
  pow_pos_n<1>(xs, kappa, r);
  return r[0];
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
