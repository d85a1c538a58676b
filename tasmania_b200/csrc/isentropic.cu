// isentropic.cu -- stand-alone (per-stencil) kernels of the isentropic core:
//   K1 step_forward_euler, K2 step_forward_euler_momentum  (prognostics/utils.py:L43-L204)
//   K3 montgomery / diagnostic_variables / height / density_and_temperature
//      (isentropic/dynamics/diagnostics.py:L319-L570)
// These are the drop-in twins of the reference's registry entries; the benchmark path uses
// the fused kernels of isentropic_fused.cu which share the point formulas.
//
// Roofline: HBM.  K1 dry: 40 B/pt, K2: 96 B/pt, K3 montgomery: 16 B/pt.
// K3 runs one thread per (i, j) column with i along the warp, so every level is a
// coalesced row access; the scans are sequential in k exactly as in the reference (no tree
// reduction) to keep the summation order.
#include <stdlib.h>
#include <string.h>

#include "stencil_math.cuh"

using namespace tb200;

namespace {

struct Tracers {
  View now[3], in[3], out[3], tnd[3];
  int n;
};

template <int SCHEME>
int run_k1(View s_now, View s_int, View s_new, View u, View v, View s_tnd, Tracers tr,
           double dt, double dx, double dy, const int32_t o[3], const int32_t d[3],
           cudaStream_t st) {
  const int i0 = o[0], j0 = o[1], k0 = o[2];
  const FluxConst fc = make_flux_const(dx, dy);
  return launch_box("step_forward_euler", d, st, [=] __device__(int i, int j, int k) {
    i += i0; j += j0; k += k0;
    const FaceVel w = face_velocities<SCHEME>(u, v, i, j, k, fc);
    // utils.py:L95-L99
    const double div = flux_divergence<SCHEME>(w, s_int, i, j, k, fc);
    const double tnd = s_tnd.ok() ? s_tnd(i, j, k) : 0.0;
    s_new(i, j, k) = s_now(i, j, k) - dt * (div - tnd);
    // utils.py:L101-L134
    for (int t = 0; t < tr.n; ++t) {
      const double dq = flux_divergence<SCHEME>(w, tr.in[t], i, j, k, fc);
      const double src = tr.tnd[t].ok() ? s_int(i, j, k) * tr.tnd[t](i, j, k) : 0.0;
      tr.out[t](i, j, k) = tr.now[t](i, j, k) - dt * (dq - src);
    }
  });
}

template <int SCHEME>
int run_k2(View s_now, View s_new, View u, View v, View su_now, View su_int, View su_new,
           View sv_now, View sv_int, View sv_new, View mtg_now, View mtg_new, View su_tnd,
           View sv_tnd, double dt, double dx, double dy, double eps, const int32_t o[3],
           const int32_t d[3], cudaStream_t st) {
  const int i0 = o[0], j0 = o[1], k0 = o[2];
  const FluxConst fc = make_flux_const(dx, dy);
  const CDiv two_dx = make_cdiv(2.0 * dx), two_dy = make_cdiv(2.0 * dy);
  return launch_box("step_forward_euler_momentum", d, st, [=] __device__(int i, int j, int k) {
    i += i0; j += j0; k += k0;
    const FaceVel w = face_velocities<SCHEME>(u, v, i, j, k, fc);
    const double sn = s_now(i, j, k), sw = s_new(i, j, k);
    // utils.py:L191-L197
    {
      const double div = flux_divergence<SCHEME>(w, su_int, i, j, k, fc);
      const double pg_now =
          (1.0 - eps) * sn * (mtg_now(i + 1, j, k) - mtg_now(i - 1, j, k)) / two_dx;
      const double pg_new = eps * sw * (mtg_new(i + 1, j, k) - mtg_new(i - 1, j, k)) / two_dx;
      const double tnd = su_tnd.ok() ? su_tnd(i, j, k) : 0.0;
      su_new(i, j, k) = su_now(i, j, k) - dt * (div + pg_now + pg_new - tnd);
    }
    // utils.py:L198-L204
    {
      const double div = flux_divergence<SCHEME>(w, sv_int, i, j, k, fc);
      const double pg_now =
          (1.0 - eps) * sn * (mtg_now(i, j + 1, k) - mtg_now(i, j - 1, k)) / two_dy;
      const double pg_new = eps * sw * (mtg_new(i, j + 1, k) - mtg_new(i, j - 1, k)) / two_dy;
      const double tnd = sv_tnd.ok() ? sv_tnd(i, j, k) : 0.0;
      sv_new(i, j, k) = sv_now(i, j, k) - dt * (div + pg_now + pg_new - tnd);
    }
  });
}

int flux_extent(int scheme) {
  switch (scheme) {
    case TB200_FLUX_UPWIND:
    case TB200_FLUX_CENTERED: return 1;
    case TB200_FLUX_THIRD_ORDER_UPWIND: return 2;
    case TB200_FLUX_FIFTH_ORDER_UPWIND: return 3;
    default: return -1;
  }
}

}  // namespace

#define DISPATCH_SCHEME(scheme, CALL)                                              \
  switch (scheme) {                                                                \
    case TB200_FLUX_UPWIND: return CALL(TB200_FLUX_UPWIND);                        \
    case TB200_FLUX_CENTERED: return CALL(TB200_FLUX_CENTERED);                    \
    case TB200_FLUX_THIRD_ORDER_UPWIND: return CALL(TB200_FLUX_THIRD_ORDER_UPWIND); \
    default: return CALL(TB200_FLUX_FIFTH_ORDER_UPWIND);                           \
  }

extern "C" int tb200_step_forward_euler(
    int flux_scheme, const tb200_field *s_now, const tb200_field *s_int, tb200_field *s_new,
    const tb200_field *u_int, const tb200_field *v_int, const tb200_field *s_tnd,
    const tb200_field *const *sq_now, const tb200_field *const *sq_int,
    tb200_field *const *sq_new, const tb200_field *const *q_tnd, double dt, double dx, double dy,
    const int32_t origin[3], const int32_t domain[3], void *stream) {
  const int e = flux_extent(flux_scheme);
  TB200_REQUIRE(e > 0, "step_forward_euler: unknown flux scheme %d", flux_scheme);
  View vs_now = view(s_now), vs_int = view(s_int), vs_new = view(s_new);
  View vu = view(u_int), vv = view(v_int), vt = view(s_tnd);
  TB200_REQUIRE(box_inside(vs_now, origin, domain) && box_inside(vs_new, origin, domain),
                "step_forward_euler: s_now/s_new box outside storage");
  TB200_REQUIRE(box_inside(vs_int, origin, domain, e, e, e, e),
                "step_forward_euler: s_int box + extent %d outside storage", e);
  TB200_REQUIRE(box_inside(vu, origin, domain, 0, 1) && box_inside(vv, origin, domain, 0, 0, 0, 1),
                "step_forward_euler: u_int/v_int box outside storage");
  TB200_REQUIRE(!vt.ok() || box_inside(vt, origin, domain), "step_forward_euler: s_tnd box");
  TB200_REQUIRE(vs_int.p != vs_new.p, "step_forward_euler: s_int and s_new must not alias");
  Tracers tr{};
  if (sq_now != nullptr) {
    TB200_REQUIRE(sq_int != nullptr && sq_new != nullptr, "step_forward_euler: tracer arrays");
    tr.n = 3;
    for (int t = 0; t < 3; ++t) {
      tr.now[t] = view(sq_now[t]);
      tr.in[t] = view(sq_int[t]);
      tr.out[t] = view(sq_new[t]);
      tr.tnd[t] = q_tnd ? view(q_tnd[t]) : View{};
      TB200_REQUIRE(box_inside(tr.now[t], origin, domain) && box_inside(tr.out[t], origin, domain) &&
                        box_inside(tr.in[t], origin, domain, e, e, e, e),
                    "step_forward_euler: tracer %d box outside storage", t);
      TB200_REQUIRE(!tr.tnd[t].ok() || box_inside(tr.tnd[t], origin, domain),
                    "step_forward_euler: tracer tendency %d box outside storage", t);
      TB200_REQUIRE(tr.in[t].p != tr.out[t].p, "step_forward_euler: tracer in/out alias");
    }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define K1_CALL(S) run_k1<S>(vs_now, vs_int, vs_new, vu, vv, vt, tr, dt, dx, dy, origin, domain, st)
  DISPATCH_SCHEME(flux_scheme, K1_CALL)
}

extern "C" int tb200_step_forward_euler_momentum(
    int flux_scheme, const tb200_field *s_now, const tb200_field *s_new,
    const tb200_field *u_int, const tb200_field *v_int, const tb200_field *su_now,
    const tb200_field *su_int, tb200_field *su_new, const tb200_field *sv_now,
    const tb200_field *sv_int, tb200_field *sv_new, const tb200_field *mtg_now,
    const tb200_field *mtg_new, const tb200_field *su_tnd, const tb200_field *sv_tnd, double dt,
    double dx, double dy, double eps, const int32_t origin[3], const int32_t domain[3],
    void *stream) {
  const int e = flux_extent(flux_scheme);
  TB200_REQUIRE(e > 0, "step_forward_euler_momentum: unknown flux scheme %d", flux_scheme);
  View a = view(s_now), b = view(s_new), vu = view(u_int), vv = view(v_int);
  View un = view(su_now), ui = view(su_int), uo = view(su_new);
  View vn = view(sv_now), vi = view(sv_int), vo = view(sv_new);
  View mn = view(mtg_now), mw = view(mtg_new), tu = view(su_tnd), tv = view(sv_tnd);
  TB200_REQUIRE(box_inside(a, origin, domain) && box_inside(b, origin, domain) &&
                    box_inside(un, origin, domain) && box_inside(uo, origin, domain) &&
                    box_inside(vn, origin, domain) && box_inside(vo, origin, domain),
                "step_forward_euler_momentum: box outside storage");
  TB200_REQUIRE(box_inside(ui, origin, domain, e, e, e, e) && box_inside(vi, origin, domain, e, e, e, e),
                "step_forward_euler_momentum: su_int/sv_int box + extent outside storage");
  TB200_REQUIRE(box_inside(mn, origin, domain, 1, 1, 1, 1) && box_inside(mw, origin, domain, 1, 1, 1, 1),
                "step_forward_euler_momentum: mtg box + 1 outside storage");
  TB200_REQUIRE(box_inside(vu, origin, domain, 0, 1) && box_inside(vv, origin, domain, 0, 0, 0, 1),
                "step_forward_euler_momentum: u_int/v_int box outside storage");
  TB200_REQUIRE((!tu.ok() || box_inside(tu, origin, domain)) && (!tv.ok() || box_inside(tv, origin, domain)),
                "step_forward_euler_momentum: tendency box outside storage");
  TB200_REQUIRE(ui.p != uo.p && vi.p != vo.p, "step_forward_euler_momentum: int/new alias");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define K2_CALL(S) \
  run_k2<S>(a, b, vu, vv, un, ui, uo, vn, vi, vo, mn, mw, tu, tv, dt, dx, dy, eps, origin, domain, st)
  DISPATCH_SCHEME(flux_scheme, K2_CALL)
}

// ---------------------------------------------------------------------------- K3
// One thread per column (i along the warp: every level is a coalesced row access).
// WHAT: 0 montgomery, 1 diagnostic_variables (p, exn, mtg, h).
//
// Two sweeps, both with running per-field pointers (one add per level, arbitrary strides):
//   down: p[k] = p[k-1] + g dz s[k-1]  -- a serial chain of one multiply-add per level, no
//         libm call; p is parked in its output storage (WHAT 1) or in mtg (WHAT 0);
//   up:   exn[k] = cp (p[k]/pref)^kappa evaluated on the way up, where the calls for
//         successive levels are independent of the cheap serial sums and pipeline;
//         mtg and h are both accumulated in this single upward pass with (p, exn) of level
//         k+1 carried in registers.
// HBM traffic (WHAT 1): read s, write p, read p, write exn, mtg, h = 48 B/point.
template <int WHAT>
static int run_k3(View theta, View hs, View s, View p, View exn, View mtg, View h, double dz,
                  double pt, double theta_s, double pref, double rd, double g, double cp,
                  const int32_t o[3], const int32_t d[3], cudaStream_t st) {
  const int i0 = o[0], j0 = o[1], k0 = o[2], nk = d[2];
  const double kappa = rd / cp;
  const CDiv cpref = make_cdiv(pref);
  const double gdz = g * dz, cpg = cp * g;
  View park = WHAT == 1 ? p : mtg;  // where the pressures wait for the upward sweep
  return launch_columns("isentropic_diagnostics", d[0], d[1], st, [=] __device__(int i, int j) {
    i += i0; j += j0;
    // ---- downward pressure scan, diagnostics.py:L339-L342 / L425-L428
    const double *ps = &s(i, j, k0);
    double *pp = &park(i, j, k0);
    double pk = pt;
    if (WHAT == 1) *pp = pk;  // montgomery never needs p[k0]
    for (int k = 1; k < nk; ++k) {
      pk = pk + gdz * *ps;
      ps += s.s2;
      // WHAT 0 parks p[k] in mtg[k-1]: the upward scan consumes it just before overwriting
      if (WHAT == 1) pp += park.s2;
      *pp = pk;
      if (WHAT == 0) pp += park.s2;
    }
    // ---- upward sweep.  pb, eb = pressure and Exner function of level k+1
    const int kt = k0 + nk - 1;
    double pb = pk;
    double eb = cp * pow_pos(pb / cpref, kappa);
    const double th_s = WHAT == 1 ? theta(i, j, kt) : theta_s;
    const double mtg_s = th_s * eb + g * hs(i, j, kt);  // L347 / L434
    double m = mtg_s + 0.5 * dz * eb;
    double *pm = &mtg(i, j, kt - 1);
    if (WHAT == 0) {
      *pm = m;
      for (int k = kt - 2; k >= k0; --k) {
        pm -= mtg.s2;
        const double e = cp * pow_pos(*pm / cpref, kappa);  // exn[k+1] from the parked p[k+1]
        m = m + dz * e;
        *pm = m;
      }
    } else {
      const double *pth = &theta(i, j, kt);
      double *pe = &exn(i, j, kt), *ph = &h(i, j, kt);
      pp = &park(i, j, kt);
      double thb = *pth;
      double hk = hs(i, j, kt);  // L354
      *pe = eb;
      *ph = hk;
      for (int k = kt - 1; k >= k0; --k) {
        pp -= park.s2; pe -= exn.s2; ph -= h.s2; pth -= theta.s2;
        const double pa = *pp;
        const double ea = cp * pow_pos(pa / cpref, kappa);
        const double tha = *pth;
        *pe = ea;
        // mtg[k] = mtg[k+1] + dz exn[k+1]; the first one (k = kt-1) is m itself
        if (k < kt - 1) m = m + dz * eb;
        *pm = m;
        pm -= mtg.s2;
        // L356-L360
        hk = hk - rd * (tha * ea + thb * eb) * (pa - pb) / (cpg * (pa + pb));
        *ph = hk;
        pb = pa; eb = ea; thb = tha;
      }
    }
  });
}

// diagnostic_variables for columns of at most 64 layers: the column of pressures stays in
// registers between the two sweeps (read s, write p, exn, mtg, h = 40 B/point, the algorithmic
// minimum; the generic kernel above parks p in memory and reads it back), and the Exner
// function is evaluated eight interfaces per call with interleaved chains (pow_pos_n<8>).
// Same operations in the same order as run_k3<1>, hence the same bits.
struct E8 {
  double v[8];
};
__device__ __noinline__ E8 exner8_diag(E8 x, double kappa, double cp) {
  E8 r;
  pow_pos_n<8>(x.v, kappa, r.v);
#pragma unroll
  for (int n = 0; n < 8; ++n) r.v[n] = cp * r.v[n];
  return r;
}

struct DiagArgs {
  View theta, hs, s, p, exn, mtg, h;
  double dz, pt, rd, g, cp, kappa;
  CDiv cpref;
  int i0, j0, k0, di, dj, n;  // n = number of layers = interface levels - 1
};

template <bool EXACT>
__global__ void __launch_bounds__(128, 2) diag_column_kernel(const DiagArgs a) {
  constexpr int NC = 64;
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  const int jj = blockIdx.y * blockDim.y + threadIdx.y;
  if (ii >= a.di || jj >= a.dj) return;
  const int i = ii + a.i0, j = jj + a.j0, k0 = a.k0;
  const int n = EXACT ? NC : a.n;
  const double gdz = a.g * a.dz, cpg = a.cp * a.g;
  double pr[NC];  // pr[l] = pressure at interface k0 + l + 1
  {
    const double *ps = &a.s(i, j, k0);
#pragma unroll
    for (int l = 0; l < NC; ++l) pr[l] = l < n ? __ldg(ps + l * a.s.s2) : 0.0;
  }
  // ---- downward pressure scan, diagnostics.py:L339-L342
  {
    double *pp = &a.p(i, j, k0);
    double pk = a.pt;
    *pp = pk;
#pragma unroll
    for (int l = 0; l < NC; ++l) {
      if (l < n) {
        pk = pk + gdz * pr[l];
        pr[l] = pk;
        pp[(l + 1) * a.p.s2] = pk;
      }
    }
  }
  // ---- upward sweep, diagnostics.py:L345-L360; (pb, eb, thb) belong to interface k + 1
  const int kt = k0 + n;
  double pb = 0.0, eb = 0.0, thb = 0.0, m = 0.0, hk = 0.0;
  double *pe = &a.exn(i, j, k0), *pm = &a.mtg(i, j, k0), *ph = &a.h(i, j, k0);
  const double *pth = &a.theta(i, j, k0);
  // One interface of the upward sweep: x = rd (th_a e_a + th_b e_b) (p_a - p_b) / (cp g (p_a + p_b))
  // is the height increment (L356-L360); the increments of a group of eight interfaces are
  // evaluated side by side (eight independent IEEE divisions in flight), only the running sums
  // (h and the Montgomery potential) are serial.
  auto increment = [&](double tha, double ea, double pa, double thb_, double eb_, double pb_) {
    return a.rd * (tha * ea + thb_ * eb_) * (pa - pb_) / (cpg * (pa + pb_));
  };
  auto interface = [&](int l1, double pa, double ea, double tha, double x) {  // interface k0 + l1 < k0 + n
    pe[l1 * a.exn.s2] = ea;
    if (l1 < n - 1) m = m + a.dz * eb;  // mtg[k] = mtg[k+1] + dz exn[k+1]
    pm[l1 * a.mtg.s2] = m;
    hk = hk - x;
    ph[l1 * a.h.s2] = hk;
    pb = pa; eb = ea; thb = tha;
  };
#pragma unroll
  for (int q = NC / 8 - 1; q >= 0; --q) {
    if (8 * q < n) {
      E8 x;
#pragma unroll
      for (int c = 0; c < 8; ++c) x.v[c] = 8 * q + c < n ? pr[8 * q + c] / a.cpref : 1.0;
      const E8 ex = exner8_diag(x, a.kappa, a.cp);
      // theta at the interfaces 8q+1 .. 8q+8 (those of pr[8q .. 8q+7])
      double th[8], inc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) th[c] = 8 * q + c < n ? __ldg(pth + (8 * q + c + 1) * a.theta.s2) : 0.0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int l = 8 * q + c;
        if (l < n - 1) {
          // the interface below: the next entry of this group, or the one carried from group q+1
          const double thb_ = c < 7 ? th[c + 1] : thb, eb_ = c < 7 ? ex.v[c + 1] : eb;
          const double pb_ = c < 7 ? pr[l + 1] : pb;
          inc[c] = increment(th[c], ex.v[c], pr[l], thb_, eb_, pb_);
        } else {
          inc[c] = 0.0;
        }
      }
#pragma unroll
      for (int c = 7; c >= 0; --c) {
        const int l = 8 * q + c;  // interface k0 + l + 1
        if (l == n - 1) {         // the lowest interface starts the sweep
          pb = pr[l];
          eb = ex.v[c];
          thb = th[c];
          hk = a.hs(i, j, kt);
          const double mtg_s = thb * eb + a.g * hk;  // L347
          m = mtg_s + 0.5 * a.dz * eb;
          pe[n * a.exn.s2] = eb;
          ph[n * a.h.s2] = hk;  // L354
        } else if (l < n - 1) {
          interface(l + 1, pr[l], ex.v[c], th[c], inc[c]);
        }
      }
    }
  }
  {
    const double tha = __ldg(pth), ea = a.cp * pow_pos(a.pt / a.cpref, a.kappa);
    interface(0, a.pt, ea, tha, increment(tha, ea, a.pt, thb, eb, pb));
  }
}

// diagnostic_variables for SMALL grids (same idea as stage_b_coop_kernel of the fused stage): one
// thread per column leaves a 161 x 161 grid with 0.7 warps per scheduler walking through ~5000
// dependent instructions (68 us where the traffic would allow 5).  A block of 32 columns x 8
// threads keeps only what is serial serial -- the pressure prefix sum and the two suffix sums
// (Montgomery potential, height), one add per level each -- and spreads the divisions, the Exner
// powers (eight interfaces per thread: exactly one exner8_diag call) and the height increments
// over all 256 threads through shared memory.  Every value is produced by the operation sequence
// of diag_column_kernel, hence the same bits.
__global__ void __launch_bounds__(256) diag_coop_kernel(const DiagArgs a) {
  constexpr int NC = 64;
  __shared__ double P[NC + 1][33];   // pressure at interface l; later the heights
  __shared__ double EX[NC + 1][33];  // Exner function at interface l; later the Montgomery potential
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ii = blockIdx.x * 32 + tx, jj = blockIdx.y;
  const bool live = ii < a.di;
  const int i = (live ? ii : a.di - 1) + a.i0, j = jj + a.j0, k0 = a.k0;
  const int n = a.n;
  const double gdz = a.g * a.dz, cpg = a.cp * a.g;
  // increments of the pressure sum (rows of 32 columns: coalesced)
  {
    const double *ps = &a.s(i, j, k0);
#pragma unroll
    for (int l = ty; l < NC; l += 8) P[l + 1][tx] = l < n ? gdz * __ldg(ps + l * a.s.s2) : 0.0;
  }
  __syncthreads();
  if (ty == 0) {  // diagnostics.py:L339-L342
    double pk = a.pt;
    P[0][tx] = pk;
#pragma unroll
    for (int l = 0; l < NC; ++l) {
      if (l < n) {
        pk = pk + P[l + 1][tx];
        P[l + 1][tx] = pk;
      }
    }
  }
  __syncthreads();
  {  // Exner function of the interfaces 8 ty + 1 .. 8 ty + 8 (and of the top one), p and exn out
    E8 x;
#pragma unroll
    for (int c = 0; c < 8; ++c) x.v[c] = 8 * ty + c < n ? P[8 * ty + c + 1][tx] / a.cpref : 1.0;
    if (8 * ty < n) {
      const E8 ex = exner8_diag(x, a.kappa, a.cp);
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (8 * ty + c < n) EX[8 * ty + c + 1][tx] = ex.v[c];
    }
    if (ty == 7) EX[0][tx] = a.cp * pow_pos(a.pt / a.cpref, a.kappa);
  }
  __syncthreads();
  double inc[8];  // height increments of the interfaces 8 ty + c (diagnostics.py:L356-L360)
  {
    const double *pth = &a.theta(i, j, k0);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int l1 = 8 * ty + c;
      inc[c] = 0.0;
      if (l1 < n) {
        const double tha = __ldg(pth + l1 * a.theta.s2), thb = __ldg(pth + (l1 + 1) * a.theta.s2);
        const double pa = P[l1][tx], pb = P[l1 + 1][tx];
        inc[c] = a.rd * (tha * EX[l1][tx] + thb * EX[l1 + 1][tx]) * (pa - pb) / (cpg * (pa + pb));
      }
    }
    if (live) {  // p and exn are final: store them (rows of 32 columns)
      double *pp = &a.p(i, j, k0), *pe = &a.exn(i, j, k0);
#pragma unroll
      for (int l = ty; l < NC + 1; l += 8) {
        if (l <= n) {
          pp[l * a.p.s2] = P[l][tx];
          pe[l * a.exn.s2] = EX[l][tx];
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (8 * ty + c < n) P[8 * ty + c][tx] = inc[c];  // the pressures are no longer needed
  __syncthreads();
  if (ty == 0) {  // upward sums, diagnostics.py:L345-L360
    const double eb_s = EX[n][tx];
    double hk = a.hs(i, j, k0 + n);
    const double thb = __ldg(&a.theta(i, j, k0) + n * a.theta.s2);
    const double mtg_s = thb * eb_s + a.g * hk;  // L347
    double m = mtg_s + 0.5 * a.dz * eb_s;
    P[n][tx] = hk;  // L354
    for (int l1 = n - 1; l1 >= 0; --l1) {
      if (l1 < n - 1) m = m + a.dz * EX[l1 + 1][tx];  // mtg[k] = mtg[k+1] + dz exn[k+1]
      EX[l1 + 1][tx] = m;                               // parked one slot up (exn[l1+1] is spent)
      hk = hk - P[l1][tx];
      P[l1][tx] = hk;
    }
  }
  __syncthreads();
  if (live) {
    double *pm = &a.mtg(i, j, k0), *ph = &a.h(i, j, k0);
#pragma unroll
    for (int l = ty; l < NC + 1; l += 8) {
      if (l <= n) ph[l * a.h.s2] = P[l][tx];
      if (l < n) pm[l * a.mtg.s2] = EX[l + 1][tx];
    }
  }
}

extern "C" int tb200_montgomery(const tb200_field *in_hs, const tb200_field *in_s,
                                tb200_field *inout_mtg, double dz, double pt, double theta_s,
                                const double constants[4], const int32_t origin[3],
                                const int32_t domain[3], void *stream) {
  View hs = view(in_hs), s = view(in_s), mtg = view(inout_mtg);
  TB200_REQUIRE(domain[2] >= 2, "montgomery: needs at least two interface levels");
  TB200_REQUIRE(box_inside(hs, origin, domain) && box_inside(s, origin, domain) &&
                    box_inside(mtg, origin, domain),
                "montgomery: box outside storage");
  TB200_REQUIRE(s.p != mtg.p, "montgomery: in_s and inout_mtg must not alias");
  return run_k3<0>(View{}, hs, s, View{}, View{}, mtg, View{}, dz, pt, theta_s, constants[0],
                   constants[1], constants[2], constants[3], origin, domain,
                   static_cast<cudaStream_t>(stream));
}

extern "C" int tb200_diagnostic_variables(const tb200_field *in_theta, const tb200_field *in_hs,
                                          const tb200_field *in_s, tb200_field *inout_p,
                                          tb200_field *out_exn, tb200_field *inout_mtg,
                                          tb200_field *inout_h, double dz, double pt,
                                          const double constants[4], const int32_t origin[3],
                                          const int32_t domain[3], void *stream) {
  View th = view(in_theta), hs = view(in_hs), s = view(in_s);
  View p = view(inout_p), exn = view(out_exn), mtg = view(inout_mtg), h = view(inout_h);
  TB200_REQUIRE(domain[2] >= 2, "diagnostic_variables: needs at least two interface levels");
  TB200_REQUIRE(box_inside(th, origin, domain) && box_inside(hs, origin, domain) &&
                    box_inside(s, origin, domain) && box_inside(p, origin, domain) &&
                    box_inside(exn, origin, domain) && box_inside(mtg, origin, domain) &&
                    box_inside(h, origin, domain),
                "diagnostic_variables: box outside storage");
  if (domain[2] - 1 <= 64 && domain[0] > 0 && domain[1] > 0) {  // register-resident columns
    DiagArgs a{th, hs, s, p, exn, mtg, h, dz, pt, constants[1], constants[2], constants[3],
               constants[1] / constants[3], make_cdiv(constants[0]), origin[0], origin[1], origin[2],
               domain[0], domain[1], domain[2] - 1};
    // fewer than four warps per scheduler with one thread per column: the cooperative kernel
    // (TB200_DIAG_IMPL=column|coop forces one)
    static int forced = -1;
    if (forced < 0) {
      const char *e = getenv("TB200_DIAG_IMPL");
      forced = e == nullptr ? 0 : strcmp(e, "column") == 0 ? 1 : strcmp(e, "coop") == 0 ? 2 : 0;
    }
    const bool coop = forced ? forced == 2 : (long long)domain[0] * domain[1] < 148LL * 4 * 4 * 32;
    if (coop) {
      dim3 block(32, 8, 1);
      dim3 grid((domain[0] + 31) / 32, domain[1], 1);
      diag_coop_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(a);
      return check_launch("diagnostic_variables(coop)");
    }
    dim3 block(32, 4, 1);
    dim3 grid((domain[0] + 31) / 32, (domain[1] + 3) / 4, 1);
    if (a.n == 64)
      diag_column_kernel<true><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(a);
    else
      diag_column_kernel<false><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("diagnostic_variables");
  }
  return run_k3<1>(th, hs, s, p, exn, mtg, h, dz, pt, 0.0, constants[0], constants[1],
                   constants[2], constants[3], origin, domain, static_cast<cudaStream_t>(stream));
}

// height: p and exn of the column are needed together on the way up.  The kernel walks down
// once storing p[k] into inout_h[k] (h is pure output), then walks up keeping
// (p[k+1], exn[k+1]) in registers and recomputing exn[k] = cp (p[k]/pref)^kappa from the
// parked p[k] -- the same expression on the same operand, hence the same bits.
extern "C" int tb200_height(const tb200_field *in_theta, const tb200_field *in_hs,
                            const tb200_field *in_s, tb200_field *inout_h, double dz, double pt,
                            const double constants[4], const int32_t origin[3],
                            const int32_t domain[3], void *stream) {
  View th = view(in_theta), hs = view(in_hs), s = view(in_s), h = view(inout_h);
  TB200_REQUIRE(domain[2] >= 2, "height: needs at least two interface levels");
  TB200_REQUIRE(box_inside(th, origin, domain) && box_inside(hs, origin, domain) &&
                    box_inside(s, origin, domain) && box_inside(h, origin, domain),
                "height: box outside storage");
  TB200_REQUIRE(hs.p != h.p && s.p != h.p, "height: inputs must not alias inout_h");
  const double pref = constants[0], rd = constants[1], g = constants[2], cp = constants[3];
  const double kappa = rd / cp;
  const CDiv cpref = make_cdiv(pref);
  const int i0 = origin[0], j0 = origin[1], k0 = origin[2], k1 = origin[2] + domain[2];
  return launch_columns("height", domain[0], domain[1], static_cast<cudaStream_t>(stream),
                        [=] __device__(int i, int j) {
                          i += i0; j += j0;
                          double pk = pt;
                          for (int k = k0; k < k1; ++k) {
                            if (k > k0) pk = pk + g * dz * s(i, j, k - 1);
                            if (k < k1 - 1) h(i, j, k) = pk;  // park p[k]
                          }
                          double pb = pk;  // p[k1-1]
                          double eb = cp * pow_pos(pb / cpref, kappa);
                          double hk = hs(i, j, k1 - 1);
                          h(i, j, k1 - 1) = hk;
                          for (int k = k1 - 2; k >= k0; --k) {
                            const double pa = h(i, j, k);
                            const double ea = cp * pow_pos(pa / cpref, kappa);
                            hk = hk - rd * (th(i, j, k) * ea + th(i, j, k + 1) * eb) * (pa - pb) /
                                          (cp * g * (pa + pb));
                            h(i, j, k) = hk;
                            pb = pa;
                            eb = ea;
                          }
                        });
}

extern "C" int tb200_density_and_temperature(const tb200_field *in_theta,
                                             const tb200_field *in_s, const tb200_field *in_exn,
                                             const tb200_field *in_h, tb200_field *out_rho,
                                             tb200_field *out_t, double cp,
                                             const int32_t origin[3], const int32_t domain[3],
                                             void *stream) {
  View th = view(in_theta), s = view(in_s), exn = view(in_exn), h = view(in_h);
  View rho = view(out_rho), t = view(out_t);
  TB200_REQUIRE(box_inside(th, origin, domain, 0, 0, 0, 0, 0, 1) &&
                    box_inside(exn, origin, domain, 0, 0, 0, 0, 0, 1) &&
                    box_inside(h, origin, domain, 0, 0, 0, 0, 0, 1) &&
                    box_inside(s, origin, domain) && box_inside(rho, origin, domain) &&
                    box_inside(t, origin, domain),
                "density_and_temperature: box outside storage");
  const int i0 = origin[0], j0 = origin[1], k0 = origin[2];
  // diagnostics.py:L559-L570
  return launch_box("density_and_temperature", domain, static_cast<cudaStream_t>(stream),
                    [=] __device__(int i, int j, int k) {
                      i += i0; j += j0; k += k0;
                      const double ta = th(i, j, k), tb = th(i, j, k + 1);
                      rho(i, j, k) = s(i, j, k) * (ta - tb) / (h(i, j, k) - h(i, j, k + 1));
                      t(i, j, k) = 0.5 / cp * (ta * exn(i, j, k) + tb * exn(i, j, k + 1));
                    });
}
