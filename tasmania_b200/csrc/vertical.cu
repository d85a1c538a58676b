// vertical.cu -- vertical advection of the isentropic model.
//
// Reference (numpy definitions):
//   src/tasmania/isentropic/physics/vertical_advection.py:L271-L386   IsentropicVerticalAdvection
//   src/tasmania/isentropic/dynamics/subclasses/minimal_vertical_fluxes/upwind.py:L31-L33,
//   centered.py:L28-L30, third_order_upwind.py:L31-L38, fifth_order_upwind.py:L31-L42
//
// tendency[k] = (F[k+1] - F[k]) / dz for k in [k0 + e, k0 + nk - e), F[K] the flux through
// interface K (between levels K-1 and K) of s, su, sv (and of s q for the water species, whose
// tendency is further divided by s); zero elsewhere.  The vertical velocity w either lives on
// the interfaces or is averaged onto them from the main levels.  One thread per point, i along
// the warp; a thread evaluates the two fluxes of its level (each interface flux is evaluated by
// the two levels it separates: the kernel stays a pure stream of 4 + 3 (7 + 6 moist) fields).
// The reference's set_output works on WHOLE storages (generics.py:L38-L40): with `overwrite` the
// output is zero outside the computed levels -- and outside the box -- so the kernel covers the
// full output storage.  Operation order is the reference's: bit-identical results.
#include <stdlib.h>
#include <string.h>

#include "stencil_math.cuh"

using namespace tb200;

namespace {

template <int SCHEME>
struct VFlux;

template <>
struct VFlux<TB200_FLUX_UPWIND> {  // upwind.py:L31-L33
  static constexpr int extent = 1;
  template <class Phi>
  __device__ __forceinline__ static double eval(double w, Phi phi, int K, const FluxConst &) {
    return w * (w > 0.0 ? phi(K) : phi(K - 1));
  }
};
template <>
struct VFlux<TB200_FLUX_CENTERED> {  // centered.py:L28-L30
  static constexpr int extent = 1;
  template <class Phi>
  __device__ __forceinline__ static double eval(double w, Phi phi, int K, const FluxConst &) {
    return w * 0.5 * (phi(K) + phi(K - 1));
  }
};
template <>
struct VFlux<TB200_FLUX_THIRD_ORDER_UPWIND> {  // third_order_upwind.py:L31-L38
  static constexpr int extent = 2;
  template <class Phi>
  __device__ __forceinline__ static double eval(double w, Phi phi, int K, const FluxConst &c) {
    const double wq = w / c.c12;  // |w| / 12 == |w / 12| bit for bit
    const double m2 = phi(K - 2), m1 = phi(K - 1), p0 = phi(K), p1 = phi(K + 1);
    return wq * (7.0 * (m1 + p0) - (m2 + p1)) - fabs(wq) * (3.0 * (m1 - p0) - (m2 - p1));
  }
};
template <>
struct VFlux<TB200_FLUX_FIFTH_ORDER_UPWIND> {  // fifth_order_upwind.py:L31-L42
  static constexpr int extent = 3;
  template <class Phi>
  __device__ __forceinline__ static double eval(double w, Phi phi, int K, const FluxConst &c) {
    const double wq = w / c.c60;
    const double m3 = phi(K - 3), m2 = phi(K - 2), m1 = phi(K - 1);
    const double p0 = phi(K), p1 = phi(K + 1), p2 = phi(K + 2);
    return wq * (37.0 * (m1 + p0) - 8.0 * (m2 + p1) + (m3 + p2)) -
           fabs(wq) * (10.0 * (m1 - p0) - 5.0 * (m2 - p1) + (m3 - p2));
  }
};

struct VAdvArgs {
  View w, s, su, sv, qv, qc, qr;
  View out[6];   // s, su, sv, qv, qc, qr
  View base[6];  // stepped mode: out = base + factor * tendency (one stage of a tendency stepper)
  double factor;
  bool ow[6];
  int nout;      // 3 dry, 6 moist
  bool staggered;
  double dz;
  CDiv cdz;  // division by dz as a correctly rounded reciprocal-multiply (same bits as `/`)
  FluxConst fc;
  int i0, j0, k0, di, dj, dk;
};

template <int SCHEME, bool STEP>
__global__ void __launch_bounds__(256) vadv_kernel(const VAdvArgs a, int n0, int n1, int n2) {
  using F = VFlux<SCHEME>;
  constexpr int E = F::extent;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= n0 || j >= n1) return;
  const bool col_in = i >= a.i0 && i < a.i0 + a.di && j >= a.j0 && j < a.j0 + a.dj;
  for (int k = blockIdx.z; k < n2; k += gridDim.z) {
    const bool inside = col_in && k >= a.k0 + E && k < a.k0 + a.dk - E;
    double tnd[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (inside) {
      // vertical velocity at the interfaces k and k+1 (vertical_advection.py:L308-L320); both lie
      // strictly inside the column, where the averaged velocity is defined
      double w0, w1;
      if (a.staggered) {
        w0 = a.w.ld(i, j, k);
        w1 = a.w.ld(i, j, k + 1);
      } else {
        const double wm = a.w.ld(i, j, k - 1), wc = a.w.ld(i, j, k), wp = a.w.ld(i, j, k + 1);
        w0 = 0.5 * (wc + wm);
        w1 = 0.5 * (wp + wc);
      }
      const View *dry[3] = {&a.s, &a.su, &a.sv};
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        const View &v = *dry[f];
        auto phi = [&](int kk) { return v.ld(i, j, kk); };
        tnd[f] = (F::eval(w1, phi, k + 1, a.fc) - F::eval(w0, phi, k, a.fc)) / a.cdz;  // L341-L353
      }
      if (a.nout == 6) {
        const View *q[3] = {&a.qv, &a.qc, &a.qr};
        const double sdz = a.s.ld(i, j, k) * a.dz;
#pragma unroll
        for (int f = 0; f < 3; ++f) {
          const View &v = *q[f];
          auto phi = [&](int kk) { return a.s.ld(i, j, kk) * v.ld(i, j, kk); };  // L323-L329
          // (qdiv: same bits as `/`; the flux difference of a water constituent is ZERO wherever there
          // is no cloud or rain, and the compiler's division takes its slow path for a zero numerator)
          tnd[3 + f] = qdiv(F::eval(w1, phi, k + 1, a.fc) - F::eval(w0, phi, k, a.fc), sdz);  // L366-L385
        }
      }
    }
    for (int f = 0; f < a.nout; ++f) {
      double &o = a.out[f](i, j, k);
      if (STEP)  // DataArrayDictOperator.fma of the stepper stage, math.py:L59-L63, same storage box
        o = a.base[f].ld(i, j, k) + a.factor * tnd[f];
      else
        o = a.ow[f] ? tnd[f] : o + tnd[f];  // generics.py:L38-L40, on the whole storage
    }
  }
}

// Marching variant (TB200_VADV_IMPL=march): one thread per column and chunk of KC levels, marching along k with
// the 2E+1 levels around k of every advected field in registers (the water constituents as the
// products s q the reference advects).  The kernel above evaluates every interface flux twice
// (once from either side) and re-reads each level 2E+1 times from L1 / L2 (75 loads per point in
// the moist model: 187 us per launch at configs[2], the top kernel of its step, profiles/
// README.md round 2); here every level is read once per chunk (+ 2E warm-up levels) and every
// flux evaluated once.  Same flux functions on the same operands in the same order: same bits.
// NOT the default: measured slower (configs[2]: 4.18 against 3.89 ms per step, 640 x 640 x 64: 22.8
// against 22.2 ms) -- six windows of 2E+1 levels cost 192 registers in the moist third-order
// variant, i.e. 8 warps per SM with every level's loads exposed; the point kernel's re-reads hit L1.
constexpr int VADV_KC = 16;
template <int SCHEME, bool STEP, int NF>
__global__ void __launch_bounds__(128) vadv_march_kernel(const VAdvArgs a, int n0, int n1, int n2) {
  using F = VFlux<SCHEME>;
  constexpr int E = F::extent;
  constexpr int NWIN = 2 * E + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= n0 || j >= n1) return;
  const bool col_in = i >= a.i0 && i < a.i0 + a.di && j >= a.j0 && j < a.j0 + a.dj;
  const int kb = blockIdx.z * VADV_KC, ke = min(kb + VADV_KC, n2);
  const int klo = a.k0 + E, khi = a.k0 + a.dk - E;  // levels with a tendency: [klo, khi)
  const View *fld[6] = {&a.s, &a.su, &a.sv, &a.qv, &a.qc, &a.qr};
  // level kk of field f as the reference advects it (clamped into the storage: clamped levels are
  // only ever part of windows of levels without a tendency)
  auto level = [&](int f, int kk, double s_at) {
    const double x = fld[f]->ld(i, j, kk);
    return f < 3 ? x : s_at * x;  // vertical_advection.py:L323-L329
  };
  double win[NF][NWIN];  // win[f][m] = field f at level k - E + m
  if (col_in) {
#pragma unroll
    for (int m = 0; m < NWIN - 1; ++m) {  // levels kb - E .. kb + E - 1 go to slots 1 .. 2E (shifted below)
      const int kk = min(max(kb - E + m, 0), n2 - 1);
      const double s_at = a.s.ld(i, j, kk);
#pragma unroll
      for (int f = 0; f < NF; ++f) win[f][m + 1] = f == 0 ? s_at : level(f, kk, s_at);
    }
  }
  double flo[NF];
  bool have_lo = false;
  for (int k = kb; k < ke; ++k) {
    double tnd[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) tnd[f] = 0.0;
    if (col_in) {
      const int kk = min(k + E, n2 - 1);
      const double s_at = a.s.ld(i, j, kk);
#pragma unroll
      for (int f = 0; f < NF; ++f) {
#pragma unroll
        for (int m = 0; m < NWIN - 1; ++m) win[f][m] = win[f][m + 1];
        win[f][NWIN - 1] = f == 0 ? s_at : level(f, kk, s_at);
      }
      if (k >= klo && k < khi) {
        double w0, w1;  // vertical velocity at the interfaces k and k + 1 (vertical_advection.py:L308-L320)
        if (a.staggered) {
          w0 = a.w.ld(i, j, k);
          w1 = a.w.ld(i, j, k + 1);
        } else {
          const double wm = a.w.ld(i, j, k - 1), wc = a.w.ld(i, j, k), wp = a.w.ld(i, j, k + 1);
          w0 = 0.5 * (wc + wm);
          w1 = 0.5 * (wp + wc);
        }
        const double sdz = win[0][E] * a.dz;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          auto phi = [&](int q) { return win[f][q - k + E]; };  // q in [k - E, k + E]
          if (!have_lo) flo[f] = F::eval(w0, phi, k, a.fc);
          const double fhi = F::eval(w1, phi, k + 1, a.fc);
          if (f < 3)
            tnd[f] = (fhi - flo[f]) / a.cdz;  // L341-L353
          else
            tnd[f] = qdiv(fhi - flo[f], sdz);  // L366-L385
          flo[f] = fhi;
        }
        have_lo = true;
      } else {
        have_lo = false;
      }
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      double &o = a.out[f](i, j, k);
      if (STEP)  // DataArrayDictOperator.fma of the stepper stage, math.py:L59-L63, same storage box
        o = a.base[f].ld(i, j, k) + a.factor * tnd[f];
      else
        o = a.ow[f] ? tnd[f] : o + tnd[f];  // generics.py:L38-L40, on the whole storage
    }
  }
}

int vadv_impl() {  // TB200_VADV_IMPL=point (default) | march
  static int impl = -1;
  if (impl < 0) {
    const char *e = getenv("TB200_VADV_IMPL");
    impl = (e != nullptr && strcmp(e, "march") == 0) ? 1 : 0;
  }
  return impl;
}

template <int SCHEME>
int run_vadv(const VAdvArgs &a, bool step, cudaStream_t st) {
  const View &o = a.out[0];
  if (o.n0 <= 0 || o.n1 <= 0 || o.n2 <= 0) return TB200_OK;
  if (vadv_impl() == 1) {
    dim3 block(32, 4, 1);
    dim3 grid((o.n0 + 31) / 32, (o.n1 + 3) / 4, (o.n2 + VADV_KC - 1) / VADV_KC);
    if (a.nout == 6) {
      if (step)
        vadv_march_kernel<SCHEME, true, 6><<<grid, block, 0, st>>>(a, o.n0, o.n1, o.n2);
      else
        vadv_march_kernel<SCHEME, false, 6><<<grid, block, 0, st>>>(a, o.n0, o.n1, o.n2);
    } else {
      if (step)
        vadv_march_kernel<SCHEME, true, 3><<<grid, block, 0, st>>>(a, o.n0, o.n1, o.n2);
      else
        vadv_march_kernel<SCHEME, false, 3><<<grid, block, 0, st>>>(a, o.n0, o.n1, o.n2);
    }
    return check_launch(step ? "vertical_advection_step(march)" : "vertical_advection(march)");
  }
  dim3 block(64, 4, 1);
  dim3 grid((o.n0 + 63) / 64, (o.n1 + 3) / 4, o.n2 > 65535 ? 65535 : o.n2);
  if (step)
    vadv_kernel<SCHEME, true><<<grid, block, 0, st>>>(a, o.n0, o.n1, o.n2);
  else
    vadv_kernel<SCHEME, false><<<grid, block, 0, st>>>(a, o.n0, o.n1, o.n2);
  return check_launch(step ? "vertical_advection_step" : "vertical_advection");
}

}  // namespace

// ---------------------------------------------------------------- implicit vertical advection
// src/tasmania/isentropic/physics/implicit_vertical_advection.py:L221-L336 with the tridiagonal
// set-up and the Thomas algorithm of
// src/tasmania/framework/subclasses/subroutine_definitions/cla.py:L42-L78, L81-L108.
// Crank-Nicolson in the vertical: for every advected field phi
//   a[k] = gamma w[k-1],  b = 1,  c[k] = -gamma w[k+1],
//   d[k] = phi[k] - gamma (w[k-1] phi[k-1] - w[k+1] phi[k+1])        (first / last level: identity)
// solved per column.  The matrix depends on w only, so the forward elimination factors
// (beta[k], the multipliers) are computed once per column and shared by the three (six) fields;
// the reference recomputes them per field with the same operations, hence the same bits.  One
// thread per column, i along the warp; beta, c and the multipliers sit in thread-local arrays,
// the eliminated right-hand side in the output storage itself.
struct ImplArgs {
  View w, in[6], out[6];  // s, su, sv, qv, qc, qr
  int nfields;            // 3 dry, 6 moist
  bool staggered;
  double gamma;
  double dt;  // > 0: return the tendencies (x_new - x) / dt (the prognostic component, L793-L919)
  int i0, j0, k0, di, dj, dk;
};

template <int MAXK>
__global__ void __launch_bounds__(128) implicit_vadv_kernel(const ImplArgs a) {
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  const int jj = blockIdx.y * blockDim.y + threadIdx.y;
  if (ii >= a.di || jj >= a.dj) return;
  const int i = ii + a.i0, j = jj + a.j0, k0 = a.k0, nk = a.dk;
  const double gamma = a.gamma, mgamma = -a.gamma;
  double beta[MAXK], cc[MAXK], mult[MAXK], wm[MAXK];
  // vertical velocity on the main levels, implicit_vertical_advection.py:L246-L252
  for (int l = 0; l < nk; ++l)
    wm[l] = a.staggered ? 0.5 * (a.w.ld(i, j, k0 + l) + a.w.ld(i, j, k0 + l + 1)) : a.w.ld(i, j, k0 + l);
  // matrix and its elimination (cla.py:L98-L102, L60-L68): a[l] = gamma w[l-1], c[l] = -gamma w[l+1]
  // on the inner levels, zero on the first and the last one
  beta[0] = 1.0;
  cc[0] = 0.0;
  mult[0] = 0.0;
  for (int l = 1; l < nk; ++l) {
    const bool inner = l < nk - 1;
    const double al = inner ? gamma * wm[l - 1] : 0.0;
    cc[l] = inner ? mgamma * wm[l + 1] : 0.0;
    const double m = beta[l - 1] != 0.0 ? al / beta[l - 1] : al;
    mult[l] = m;
    beta[l] = 1.0 - m * cc[l - 1];
  }
  for (int f = 0; f < a.nfields; ++f) {
    const bool water = f >= 3;  // advected as s q, returned as (s q)_new / s_new (L255-L262, L331-L336)
    auto phi = [&](int l) {
      const double v = a.in[f].ld(i, j, k0 + l);
      return water ? a.in[0].ld(i, j, k0 + l) * v : v;
    };
    View o = a.out[f];
    // right-hand side (cla.py:L104-L108) and forward sweep (L64-L68), delta parked in the output
    double delta = phi(0);
    o(i, j, k0) = delta;
    for (int l = 1; l < nk; ++l) {
      double d;
      if (l < nk - 1)
        d = phi(l) - gamma * (wm[l - 1] * phi(l - 1) - wm[l + 1] * phi(l + 1));
      else
        d = phi(l);
      delta = d - mult[l] * delta;
      o(i, j, k0 + l) = delta;
    }
    // backward substitution (cla.py:L70-L78); b == 1 where beta == 0.  In tendency mode the
    // solved s stays in its output until the water species have been divided by it.
    const bool tend = a.dt > 0.0;
    auto result = [&](int l, double x) {
      double v = water ? x / a.out[0](i, j, k0 + l) : x;
      if (tend && f > 0) v = (v - a.in[f].ld(i, j, k0 + l)) / a.dt;  // L908-L919
      return v;
    };
    double x = beta[nk - 1] != 0.0 ? delta / beta[nk - 1] : delta / 1.0;
    o(i, j, k0 + nk - 1) = result(nk - 1, x);
    for (int l = nk - 2; l >= 0; --l) {
      const double r = o(i, j, k0 + l) - cc[l] * x;
      x = beta[l] != 0.0 ? r / beta[l] : r / 1.0;
      o(i, j, k0 + l) = result(l, x);
    }
  }
  if (a.dt > 0.0)
    for (int l = 0; l < nk; ++l)
      a.out[0](i, j, k0 + l) = (a.out[0](i, j, k0 + l) - a.in[0].ld(i, j, k0 + l)) / a.dt;  // L908
}

extern "C" int tb200_implicit_vertical_advection(
    int staggered_w, const tb200_field *in_w, const tb200_field *in_s, const tb200_field *in_su,
    const tb200_field *in_sv, tb200_field *out_s, tb200_field *out_su, tb200_field *out_sv,
    const tb200_field *in_qv, const tb200_field *in_qc, const tb200_field *in_qr,
    tb200_field *out_qv, tb200_field *out_qc, tb200_field *out_qr, double gamma, double dt_tendency,
    const int32_t origin[3], const int32_t domain[3], void *stream) {
  ImplArgs a{};
  a.w = view(in_w);
  const tb200_field *ins[6] = {in_s, in_su, in_sv, in_qv, in_qc, in_qr};
  tb200_field *outs[6] = {out_s, out_su, out_sv, out_qv, out_qc, out_qr};
  for (int f = 0; f < 6; ++f) {
    a.in[f] = view(ins[f]);
    a.out[f] = view(outs[f]);
  }
  bool moist = false;
  for (int f = 3; f < 6; ++f) moist = moist || a.in[f].ok() || a.out[f].ok();
  a.nfields = moist ? 6 : 3;
  a.staggered = staggered_w != 0;
  a.gamma = gamma;
  a.dt = dt_tendency;
  a.i0 = origin[0]; a.j0 = origin[1]; a.k0 = origin[2];
  a.di = domain[0]; a.dj = domain[1]; a.dk = domain[2];
  TB200_REQUIRE(dt_tendency >= 0.0, "implicit_vertical_advection: dt_tendency must be >= 0");
  TB200_REQUIRE(a.dk >= 2 && a.dk <= 256, "implicit_vertical_advection: 2 <= levels <= 256, got %d", a.dk);
  TB200_REQUIRE(a.staggered ? box_inside(a.w, origin, domain, 0, 0, 0, 0, 0, 1)
                            : box_inside(a.w, origin, domain),
                "implicit_vertical_advection: box outside the vertical velocity storage");
  for (int f = 0; f < a.nfields; ++f) {
    TB200_REQUIRE(box_inside(a.in[f], origin, domain) && box_inside(a.out[f], origin, domain),
                  "implicit_vertical_advection: box outside a field storage (moist needs all six)");
    for (int g = 0; g < a.nfields; ++g)
      TB200_REQUIRE(a.out[f].p != a.in[g].p && (f == g || a.out[f].p != a.out[g].p) && a.out[f].p != a.w.p,
                    "implicit_vertical_advection: outputs must not alias inputs or each other");
  }
  if (a.di <= 0 || a.dj <= 0) return TB200_OK;
  dim3 block(32, 4, 1);
  dim3 grid((a.di + 31) / 32, (a.dj + 3) / 4, 1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a.dk <= 64)
    implicit_vadv_kernel<64><<<grid, block, 0, st>>>(a);
  else
    implicit_vadv_kernel<256><<<grid, block, 0, st>>>(a);
  return check_launch("implicit_vertical_advection");
}

// ---------------------------------------------------------------- Thomas algorithm (stand-alone)
// src/tasmania/framework/subclasses/stencil_definitions/cla.py:L33-L62 (the global `thomas`
// stencil): per column, forward elimination
//   w = beta[k-1] != 0 ? a[k] / beta[k-1] : a[k];  beta[k] = b[k] - w c[k-1];  delta[k] = d[k] - w delta[k-1]
// and back substitution  x[k] = (delta[k] - c[k] x[k+1]) / (beta[k] != 0 ? beta[k] : b[k]).
// One thread per column, i along the warp (every level a coalesced row access); beta in a
// thread-local array, the eliminated right-hand side parked in the output storage (the
// reference works on deep copies of b and d).  24 + 16 B/pt: a, b, c, d read once, out written
// and re-read once.
struct ThomasArgs {
  View a, b, c, d, out;
  int i0, j0, k0, di, dj, dk;
};

template <int MAXK>
__global__ void __launch_bounds__(128) thomas_kernel(const ThomasArgs t) {
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  const int jj = blockIdx.y * blockDim.y + threadIdx.y;
  if (ii >= t.di || jj >= t.dj) return;
  const int i = ii + t.i0, j = jj + t.j0, k0 = t.k0, nk = t.dk;
  double beta[MAXK];
  // d through plain loads: out may alias it (never the read-only path for memory this kernel writes)
  double bk = t.b.ld(i, j, k0), delta = t.d(i, j, k0);
  beta[0] = bk;
  t.out(i, j, k0) = delta;
  for (int l = 1; l < nk; ++l) {
    const double al = t.a.ld(i, j, k0 + l);
    const double w = beta[l - 1] != 0.0 ? al / beta[l - 1] : al;
    bk = t.b.ld(i, j, k0 + l) - w * t.c.ld(i, j, k0 + l - 1);
    beta[l] = bk;
    delta = t.d(i, j, k0 + l) - w * delta;
    t.out(i, j, k0 + l) = delta;
  }
  double x = delta / (bk != 0.0 ? bk : t.b.ld(i, j, k0 + nk - 1));
  t.out(i, j, k0 + nk - 1) = x;
  for (int l = nk - 2; l >= 0; --l) {
    const double r = t.out(i, j, k0 + l) - t.c.ld(i, j, k0 + l) * x;
    x = r / (beta[l] != 0.0 ? beta[l] : t.b.ld(i, j, k0 + l));
    t.out(i, j, k0 + l) = x;
  }
}

extern "C" int tb200_thomas(const tb200_field *a, const tb200_field *b, const tb200_field *c,
                            const tb200_field *d, tb200_field *out, const int32_t origin[3],
                            const int32_t domain[3], void *stream) {
  ThomasArgs t{};
  t.a = view(a); t.b = view(b); t.c = view(c); t.d = view(d); t.out = view(out);
  t.i0 = origin[0]; t.j0 = origin[1]; t.k0 = origin[2];
  t.di = domain[0]; t.dj = domain[1]; t.dk = domain[2];
  TB200_REQUIRE(box_inside(t.a, origin, domain) && box_inside(t.b, origin, domain) &&
                    box_inside(t.c, origin, domain) && box_inside(t.d, origin, domain) &&
                    box_inside(t.out, origin, domain),
                "thomas: box outside a storage");
  TB200_REQUIRE(t.dk >= 1 && t.dk <= 256, "thomas: 1 <= levels <= 256, got %d", t.dk);
  TB200_REQUIRE(t.out.p != t.a.p && t.out.p != t.b.p && t.out.p != t.c.p,
                "thomas: out must not alias a, b or c");
  if (t.di <= 0 || t.dj <= 0) return TB200_OK;
  dim3 block(32, 4, 1);
  dim3 grid((t.di + 31) / 32, (t.dj + 3) / 4, 1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (t.dk <= 64)
    thomas_kernel<64><<<grid, block, 0, st>>>(t);
  else
    thomas_kernel<256><<<grid, block, 0, st>>>(t);
  return check_launch("thomas");
}

static int vadv_entry(
    int flux_scheme, int staggered_w, const tb200_field *in_w, const tb200_field *in_s,
    const tb200_field *in_su, const tb200_field *in_sv, tb200_field *out_s, tb200_field *out_su,
    tb200_field *out_sv, const tb200_field *in_qv, const tb200_field *in_qc,
    const tb200_field *in_qr, tb200_field *out_qv, tb200_field *out_qc, tb200_field *out_qr,
    double dz, uint32_t overwrite_flags, const tb200_field *const *base, double factor,
    const int32_t origin[3], const int32_t domain[3], void *stream) {
  VAdvArgs a{};
  const bool step = base != nullptr;
  a.factor = factor;
  a.w = view(in_w); a.s = view(in_s); a.su = view(in_su); a.sv = view(in_sv);
  a.qv = view(in_qv); a.qc = view(in_qc); a.qr = view(in_qr);
  a.out[0] = view(out_s); a.out[1] = view(out_su); a.out[2] = view(out_sv);
  a.out[3] = view(out_qv); a.out[4] = view(out_qc); a.out[5] = view(out_qr);
  const bool moist = a.qv.ok() || a.qc.ok() || a.qr.ok() || a.out[3].ok() || a.out[4].ok() || a.out[5].ok();
  a.nout = moist ? 6 : 3;
  a.staggered = staggered_w != 0;
  a.dz = dz;
  a.cdz = make_cdiv(dz);
  a.fc = make_flux_const(1.0, 1.0);
  a.i0 = origin[0]; a.j0 = origin[1]; a.k0 = origin[2];
  a.di = domain[0]; a.dj = domain[1]; a.dk = domain[2];
  for (int f = 0; f < 6; ++f) a.ow[f] = (overwrite_flags >> f) & 1u;
  int e = -1;
  switch (flux_scheme) {
    case TB200_FLUX_UPWIND: case TB200_FLUX_CENTERED: e = 1; break;
    case TB200_FLUX_THIRD_ORDER_UPWIND: e = 2; break;
    case TB200_FLUX_FIFTH_ORDER_UPWIND: e = 3; break;
  }
  TB200_REQUIRE(e > 0, "vertical_advection: unknown flux scheme %d", flux_scheme);
  TB200_REQUIRE(box_inside(a.s, origin, domain) && box_inside(a.su, origin, domain) &&
                    box_inside(a.sv, origin, domain),
                "vertical_advection: box outside an input storage");
  TB200_REQUIRE(a.staggered ? box_inside(a.w, origin, domain, 0, 0, 0, 0, 0, 1)
                            : box_inside(a.w, origin, domain),
                "vertical_advection: box outside the vertical velocity storage");
  if (moist)
    TB200_REQUIRE(box_inside(a.qv, origin, domain) && box_inside(a.qc, origin, domain) &&
                      box_inside(a.qr, origin, domain),
                  "vertical_advection: moist call needs in_qv, in_qc, in_qr covering the box");
  for (int f = 0; f < a.nout; ++f) {
    if (step) {
      a.base[f] = view(base[f]);
      const int32_t o0[3] = {0, 0, 0};
      const int32_t whole[3] = {a.out[f].n0, a.out[f].n1, a.out[f].n2};
      TB200_REQUIRE(box_inside(a.base[f], o0, whole),
                    "vertical_advection_step: a base storage is smaller than its output storage");
      TB200_REQUIRE(a.out[f].p != a.base[f].p, "vertical_advection_step: outputs must not alias the base fields");
    }
    TB200_REQUIRE(box_inside(a.out[f], origin, domain), "vertical_advection: box outside an output storage");
    TB200_REQUIRE(a.out[f].n0 == a.out[0].n0 && a.out[f].n1 == a.out[0].n1 && a.out[f].n2 == a.out[0].n2,
                  "vertical_advection: the output storages must share one shape");
    TB200_REQUIRE(a.out[f].p != a.s.p && a.out[f].p != a.su.p && a.out[f].p != a.sv.p &&
                      a.out[f].p != a.w.p && a.out[f].p != a.qv.p && a.out[f].p != a.qc.p &&
                      a.out[f].p != a.qr.p,
                  "vertical_advection: outputs must not alias inputs");
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (flux_scheme) {
    case TB200_FLUX_UPWIND: return run_vadv<TB200_FLUX_UPWIND>(a, step, st);
    case TB200_FLUX_CENTERED: return run_vadv<TB200_FLUX_CENTERED>(a, step, st);
    case TB200_FLUX_THIRD_ORDER_UPWIND: return run_vadv<TB200_FLUX_THIRD_ORDER_UPWIND>(a, step, st);
    default: return run_vadv<TB200_FLUX_FIFTH_ORDER_UPWIND>(a, step, st);
  }
}

extern "C" int tb200_vertical_advection(
    int flux_scheme, int staggered_w, const tb200_field *in_w, const tb200_field *in_s,
    const tb200_field *in_su, const tb200_field *in_sv, tb200_field *out_s, tb200_field *out_su,
    tb200_field *out_sv, const tb200_field *in_qv, const tb200_field *in_qc,
    const tb200_field *in_qr, tb200_field *out_qv, tb200_field *out_qc, tb200_field *out_qr,
    double dz, uint32_t overwrite_flags, const int32_t origin[3], const int32_t domain[3],
    void *stream) {
  return vadv_entry(flux_scheme, staggered_w, in_w, in_s, in_su, in_sv, out_s, out_su, out_sv, in_qv,
                    in_qc, in_qr, out_qv, out_qc, out_qr, dz, overwrite_flags, nullptr, 0.0, origin,
                    domain, stream);
}

// One stage of a tendency stepper around the vertical advection in one kernel: the tendencies of
// `in` are formed exactly as above and out[f] = base[f] + factor * tendency[f] on the whole output
// storage -- what tb200_vertical_advection followed by DataArrayDictOperator.fma
// (src/tasmania/utils/xarrayx.py:L688-L740) computes, without writing and re-reading the
// tendencies.  in / base / out: s, su, sv[, qv, qc, qr] (nfields = 3 or 6).
extern "C" int tb200_vertical_advection_step(
    int flux_scheme, int staggered_w, const tb200_field *in_w, int nfields,
    const tb200_field *const *in, const tb200_field *const *base, tb200_field *const *out, double dz,
    double factor, const int32_t origin[3], const int32_t domain[3], void *stream) {
  TB200_REQUIRE(in != nullptr && base != nullptr && out != nullptr, "vertical_advection_step: NULL argument");
  TB200_REQUIRE(nfields == 3 || nfields == 6, "vertical_advection_step: 3 or 6 fields, got %d", nfields);
  for (int f = 0; f < nfields; ++f)
    TB200_REQUIRE(in[f] != nullptr && base[f] != nullptr && out[f] != nullptr,
                  "vertical_advection_step: NULL field %d", f);
  const bool m = nfields == 6;
  return vadv_entry(flux_scheme, staggered_w, in_w, in[0], in[1], in[2], out[0], out[1], out[2],
                    m ? in[3] : nullptr, m ? in[4] : nullptr, m ? in[5] : nullptr, m ? out[3] : nullptr,
                    m ? out[4] : nullptr, m ? out[5] : nullptr, dz, 0u, base, factor, origin, domain,
                    stream);
}
