// isentropic_fused.cu -- one Runge-Kutta stage of the dry isentropic dynamical core with the
// relaxed lateral boundary, fused into three kernels (the benchmark hot path).
//
// Reference sequence per stage (src/tasmania/isentropic/dynamics/dycore.py:L641-L721 and
// subclasses/prognostics/rk3ws_si.py:L105-L234):
//   K1 step s -> irelax(s) -> montgomery(s_new) -> K2 step su, sv -> irelax(s, su, sv, u, v)
//   -> Rayleigh damping (s, su, sv) -> velocity_x / velocity_y -> outermost layers of u, v.
// The reference runs 13 full-domain passes for this; here:
//
//   kernel S   (one thread per column, marching in k)
//       s_pre = irelax(K1(...))                      -> scratch_s
//       p by the downward scan on s_pre              -> scratch_exn  (pressure at interface k+1)
//       exn = cp (p/pref)^kappa, mtg_new by the upward scan -> scratch_mtg
//   kernel MV  (one warp per 30 columns x 64 rows of a level, marching in j)
//       su, sv = irelax(K2(...)), s = irelax(s_pre), Rayleigh damping on all three,
//       u, v from the final s, su, sv, outermost faces from the reference state.
//
// The vertical scans are inherently two sweeps (pressure top-down, Montgomery bottom-up) and
// the momentum step needs mtg_new at i+-1 / j+-1, hence the kernel boundary between S and MV.
//
// HBM traffic per point and stage (8-byte words): S reads s_now, s_int, u, v, writes s_pre,
// p, re-reads p, writes mtg = 8 words; MV reads s_now, s_pre, mtg_now, mtg_new, u, v, su_now,
// su_int, sv_now, sv_int, writes s, su, sv, u, v = 15 words.  Total 23 words = 184 B against
// the algorithmic minimum of 112 B; the relaxation band and the damping
// layer add reads of the reference fields only where gamma != 0 or R != 0.  All arithmetic
// follows the reference's operation order (see stencil_math.cuh).
#include <stdlib.h>
#include <string.h>

#include "stage.cuh"

using namespace tb200;

namespace {

// ---------------------------------------------------------------- kernel S
// Unit i-stride and 32-bit element offsets (a field has < 2^31 elements) keep the address
// arithmetic off the critical path: one base pointer per field, immediate offsets along i.
template <int SCHEME>
__device__ __forceinline__ double div_unit(const FaceVel &w, const double *pc, int sj,
                                           const FluxConst &c) {
  using F = Flux<SCHEME>;
  const double fxm = F::face(w.xm, pc, 1);
  const double fxp = F::face(w.xp, pc + 1, 1);
  const double fym = F::face(w.ym, pc, sj);
  const double fyp = F::face(w.yp, pc + sj, sj);
  return (fxp - fxm) / c.dx + (fyp - fym) / c.dy;
}

template <int SCHEME>
__global__ void __launch_bounds__(128) stage_s_kernel(const StageArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= a.nx || j >= a.ny) return;
  const bool interior = i >= a.nb && i < a.nx - a.nb && j >= a.nb && j < a.ny - a.nb;
  const double gam = a.gamma.ld(i, j, 0);
  const double kappa = a.rd / a.cp;
  const double gdz = a.g * a.dz;
  using F = Flux<SCHEME>;

  // downward sweep: step s, relax, integrate the pressure; p[k+1] is parked in the exn
  // scratch (no libm call on this serial chain)
  double p = a.pt;
  {
    const double *ps_int = a.s_int.p + (i + j * a.s_int.s1);
    const double *ps_now = a.s_now.p + (i + j * a.s_now.s1);
    const double *pu = a.u_int.p + (i + j * a.u_int.s1);
    const double *pv = a.v_int.p + (i + j * a.v_int.s1);
    const double *pref_ = a.s_ref.p + (i + j * a.s_ref.s1);
    double *ps_new = a.spre.p + (i + j * a.spre.s1);
    const double *ps_old = a.s_new.p + (i + j * a.s_new.s1);
    double *pex = a.exn.p + (i + j * a.exn.s1);
    const int sj = (int)a.s_int.s1, svj = (int)a.v_int.s1;
    const long long s2 = a.s_int.s2;  // plane stride (shared by all 3-D fields)
#pragma unroll 2
    for (int k = 0; k < a.nz; ++k) {
      double v;
      if (interior && k + 2 < a.nz) {  // DRAM -> L2 two levels ahead
        prefetch_l2(ps_int + 2 * s2, 0);
        prefetch_l2(ps_int + 2 * s2 - 3 * sj, 0);
        prefetch_l2(ps_int + 2 * s2 + 3 * sj, 0);
        prefetch_l2(ps_now + 2 * s2, 0);
        prefetch_l2(pu + 2 * s2, 0);
        prefetch_l2(pv + 2 * s2, 0);
        prefetch_l2(pv + 2 * s2 + svj, 0);
      }
      if (interior) {
        const FaceVel w{F::prep(__ldg(pu), a.fc), F::prep(__ldg(pu + 1), a.fc),
                        F::prep(__ldg(pv), a.fc), F::prep(__ldg(pv + svj), a.fc)};
        const double div = div_unit<SCHEME>(w, ps_int, sj, a.fc);
        v = __ldg(ps_now) - a.dt * (div - 0.0);
      } else {
        v = gam == 1.0 ? 0.0 : *ps_old;  // untouched by K1; the relaxation below decides
      }
      if (gam != 0.0) v = relax_point(gam, v, __ldg(pref_));
      *ps_new = v;
      p = p + gdz * v;
      *pex = p;
      ps_int += a.s_int.s2; ps_now += a.s_now.s2; pu += a.u_int.s2; pv += a.v_int.s2;
      pref_ += a.s_ref.s2; ps_new += a.spre.s2; ps_old += a.s_new.s2; pex += a.exn.s2;
    }
  }
  // upward sweep, diagnostics.py:L433-L438: exn[k+1] = cp (p[k+1] / pref)^kappa from the
  // parked pressures -- the pow calls of different levels are independent of the cheap
  // serial sum, so they pipeline
  {
    const double *pex = a.exn.p + (i + j * a.exn.s1 + (long long)(a.nz - 1) * a.exn.s2);
    double *pm = a.mtg.p + (i + j * a.mtg.s1 + (long long)(a.nz - 1) * a.mtg.s2);
    const double ex_s = a.cp * pow_pos(*pex / a.cpref, kappa);
    const double mtg_s = a.theta_s * ex_s + a.g * a.hs.ld(i, j, 0);
    double m = mtg_s + 0.5 * a.dz * ex_s;
    *pm = m;
#pragma unroll 4
    for (int k = a.nz - 2; k >= 0; --k) {
      pex -= a.exn.s2;
      pm -= a.mtg.s2;
      m = m + a.dz * (a.cp * pow_pos(*pex / a.cpref, kappa));
      *pm = m;
    }
  }
}

// ---------------------------------------------------------------- kernels A and B
// Kernel S above couples two things with opposite needs: the s-step is a halo-3 horizontal
// stencil (wants (i, j) marching with shared faces and parallelism over k), the scans are serial
// in k (want one thread per column).  Splitting them:
//
//   kernel A  (one warp per 31 columns x LJ rows of ONE level, marching in j like kernel MV)
//       s_pre = irelax(K1(...)) -> scratch_s.  Every face flux is evaluated once (x faces shared
//       by shuffle, y faces carried to the next row, s_int rows in a register window), and the
//       launch has nz-fold more parallelism than one thread per column.
//   kernel B  (one thread per column)
//       reads the column of s_pre into registers in one go (nz independent loads in flight per
//       thread), runs the pressure prefix sum, the Exner function and the Montgomery suffix sum
//       out of registers and writes mtg_new: no parking of p in memory.
//
// HBM traffic: A reads s_now, s_int, u, v and writes s_pre (5 words), B reads s_pre and writes
// mtg (2 words): 7 words against the 8 of kernel S, with ~35 fewer fp64 instructions per point.
// Arithmetic and results are bit-identical to kernel S.
constexpr int A_COLS = 31;

template <int SCHEME, int LJ>
__global__ void __launch_bounds__(128, 6) stage_a_kernel(const StageArgs a) {
  using F = Flux<SCHEME>;
  constexpr int E = F::extent;
  constexpr int NW = 2 * E;
  const int lane = threadIdx.x & 31;
  const int xw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (xw * A_COLS >= a.nx) return;  // warp-uniform
  const int c = xw * A_COLS + lane;  // lanes 0..30 own a column, lane 31 lends its left face
  const int j0 = blockIdx.y * LJ;
  const int jend = min(j0 + LJ, a.ny);
  const int k = blockIdx.z;
  const int nx = a.nx, ny = a.ny, nb = a.nb;

  const bool out_lane = lane < A_COLS && c < nx;
  const bool col_int = c >= nb && c < nx - nb;
  const int cc = min(max(c, E), nx - E);  // keeps every x-offset load inside the row
  const int cm = min(c, nx - 1);          // own column (lanes beyond the domain are discarded)

  const unsigned row = (unsigned)a.s_now.s1 * 8u;
  const unsigned plane = (unsigned)k * (unsigned)a.s_now.s2 * 8u;
  const unsigned grow = (unsigned)a.gamma.s1 * 8u;

  // window for the y-face j0 (rows j0-E .. j0+E-1) and its flux
  double ws[NW];
#pragma unroll
  for (int m = 0; m < NW; ++m)
    ws[m] = ldo(a.s_int.p, plane + (unsigned)max(j0 - E + m, 0) * row + (unsigned)cc * 8u);
  unsigned o_cc = plane + (unsigned)j0 * row + (unsigned)cc * 8u;
  unsigned o_cm = plane + (unsigned)j0 * row + (unsigned)cm * 8u;
  unsigned o_g = (unsigned)j0 * grow + (unsigned)cm * 8u;
  double fy = F::eval_v(F::prep(ldo(a.v_int.p, o_cc), a.fc), ws);

  struct Row {
    double s_w, v_n, u_c, s_now, gam;
  };
  auto load_row = [&](unsigned occ, unsigned ocm, unsigned og) {
    Row L;
    L.s_w = ldo(a.s_int.p, occ + E * row);
    L.v_n = ldo(a.v_int.p, occ + row);
    L.u_c = ldo(a.u_int.p, occ);
    L.s_now = ldo(a.s_now.p, ocm);
    L.gam = ldo(a.gamma.p, og);
    return L;
  };
  Row nxt = load_row(o_cc, o_cm, o_g);
  for (int r = j0; r < jend; ++r) {
    const Row cur = nxt;
    nxt = load_row(o_cc + row, o_cm + row, o_g + grow);
    if (r + 4 < jend) {  // DRAM -> L2 a few rows ahead
      prefetch_l2(a.s_int.p, o_cc + (E + 4) * row);
      prefetch_l2(a.v_int.p, o_cc + 5 * row);
      prefetch_l2(a.u_int.p, o_cc + 4 * row);
      prefetch_l2(a.s_now.p, o_cm + 4 * row);
    }
#pragma unroll
    for (int m = 0; m < NW - 1; ++m) ws[m] = ws[m + 1];
    ws[NW - 1] = cur.s_w;
    const double fy_p = F::eval_v(F::prep(cur.v_n, a.fc), ws);
    double xs[NW];
    {
      const double *ps = ptr_at(a.s_int.p, o_cc);
#pragma unroll
      for (int m = 0; m < NW; ++m) xs[m] = m == E ? ws[E - 1] : __ldg(ps + (m - E));
    }
    const double fx = F::eval_v(F::prep(cur.u_c, a.fc), xs);
    const double fx_p = __shfl_down_sync(0xffffffffu, fx, 1);

    const bool interior = col_int && r >= nb && r < ny - nb;
    const double gam = cur.gam;
    double v;
    if (interior) {  // prognostics/utils.py:L95-L99
      const double div = (fx_p - fx) / a.fc.dx + (fy_p - fy) / a.fc.dy;
      v = cur.s_now - a.dt * (div - 0.0);
    } else {
      v = gam == 1.0 ? 0.0 : ldo(a.s_new.p, o_cm);  // untouched by K1; the relaxation decides
    }
    if (gam != 0.0) v = relax_point(gam, v, ldo(a.s_ref.p, o_cm));  // rk3ws_si.py:L184-L189
    if (out_lane) sto(a.spre.p, o_cm, v);
    fy = fy_p;
    o_cc += row; o_cm += row; o_g += grow;
  }
}

// The Exner function of four levels as ONE real call: kernel B is unrolled over the levels, so
// an inlined copy of the power per level would not fit the instruction cache, and the four
// evaluations are interleaved instruction by instruction (pow_pos_n), which gives the fp64 pipe
// the instruction-level parallelism that a single dependent polynomial chain lacks.
struct D4 {
  double a, b, c, d;
};
__device__ __noinline__ D4 exner4(D4 x, double kappa, double cp) {
  const double xs[4] = {x.a, x.b, x.c, x.d};
  double r[4];
  pow_pos_n<4>(xs, kappa, r);
  return D4{cp * r[0], cp * r[1], cp * r[2], cp * r[3]};
}
// eight levels per call: with two resident warps per scheduler (kernel B keeps a column in
// registers) eight interleaved chains cover the 8-cycle DFMA latency from within one warp
struct D8 {
  double v[8];
};
__device__ __noinline__ D8 exner8(D8 x, double kappa, double cp) {
  D8 r;
  pow_pos_n<8>(x.v, kappa, r.v);
#pragma unroll
  for (int n = 0; n < 8; ++n) r.v[n] = cp * r.v[n];
  return r;
}

// EXACT: nz == NZC, known at compile time (no per-level range checks in the unrolled code)
template <int NZC, bool EXACT>
__global__ void __launch_bounds__(128, 2) stage_b_kernel(const StageArgs a) {
  static_assert(NZC % 8 == 0, "levels are processed eight at a time");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= a.nx || j >= a.ny) return;
  const int nz = EXACT ? NZC : a.nz;
  const double kappa = a.rd / a.cp;
  const double gdz = a.g * a.dz;
  double e[NZC];
  {
    const double *ps = a.spre.p + (i + j * a.spre.s1);
    const long long s2 = a.spre.s2;
#pragma unroll
    for (int k = 0; k < NZC; ++k) e[k] = k < nz ? __ldg(ps + k * s2) : 0.0;
  }
  // pressure at interface k+1 over the reference pressure (diagnostics.py:L425-L428, L431),
  // kept in place of s_pre; levels beyond nz get 1 (their Exner value is never used)
  double p = a.pt;
#pragma unroll
  for (int k = 0; k < NZC; ++k) {
    if (k < nz) {
      p = p + gdz * e[k];
      e[k] = p / a.cpref;
    } else {
      e[k] = 1.0;
    }
  }
  // Exner function of every interface, eight levels per call
#pragma unroll
  for (int k = 0; k < NZC; k += 8) {
    if (k < nz) {
      D8 x;
#pragma unroll
      for (int n = 0; n < 8; ++n) x.v[n] = e[k + n];
      const D8 r = exner8(x, kappa, a.cp);
#pragma unroll
      for (int n = 0; n < 8; ++n) e[k + n] = r.v[n];
    }
  }
  // upward sweep, diagnostics.py:L433-L438
  double *pm = a.mtg.p + (i + j * a.mtg.s1);
  const long long m2 = a.mtg.s2;
  double m = 0.0;
#pragma unroll
  for (int k = NZC - 1; k >= 0; --k) {
    if (k == nz - 1) {
      const double ex_s = e[k];
      const double mtg_s = a.theta_s * ex_s + a.g * a.hs.ld(i, j, 0);
      m = mtg_s + 0.5 * a.dz * ex_s;
      pm[k * m2] = m;
    } else if (k < nz - 1) {
      m = m + a.dz * e[k];
      pm[k * m2] = m;
    }
  }
}

// ---------------------------------------------------------------- kernel MV
// Momentum step + second relaxation + Rayleigh damping + velocity diagnosis in one pass.
//
// Work decomposition: one WARP owns 30 consecutive columns of one k-level and marches along
// j over a strip of LJ rows.  Lanes 1..30 produce output; lane 0 re-computes the column to
// the left (its final su, s feed lane 1's u) and lane 31 only contributes its left-face flux
// to lane 30 -- so warps are fully autonomous: no shared memory, no block barrier.
//   * y direction: each lane keeps a register window of the 2e rows of su_int / sv_int the
//     next y-face needs (one new load per row instead of 2e+1) and carries the face flux
//     F_y(j+1/2) over to the next row, so every y-face is evaluated once;
//   * x direction: a lane evaluates only its LEFT face and receives the right one from
//     lane+1 by shuffle, so every x-face is evaluated once (plus 2/32 redundancy); the x
//     neighbours of the advected field come from L1 through immediate-offset loads;
//   * the final (relaxed, damped) su, sv, s of the previous row / left lane stay in
//     registers for the velocity diagnosis, hence no extra pass over the outputs.
// A strip starts by re-computing row j0-1 (1/LJ redundancy) to seed those carries.
// s_pre is read from its own scratch (written by kernel S) because halo lanes / rows read
// points that belong to other warps, which may already have stored their final s.
// Requirement: gamma == 1 on the nb outermost rings (true for the Relaxed boundary,
// relaxed.py:L209-L211), so that the stale su/sv there never matter.
constexpr int MV_COLS = 30;

// First-touch (DRAM-latency) loads of one row, issued one row ahead of their use so that a
// warp's arithmetic on row r overlaps its own memory traffic for row r+1.
struct RowLoads {
  double su_w, sv_w;  // su_int, sv_int at row r+E   (newest window entry)
  double v_n;         // v_int at row r+1            (advects through the y-face r+1)
  double u_c;         // u_int at row r              (advects through the left x-face)
  double s_pre, s_now, su_now, sv_now;  // row r
  double mn_p, mw_p;  // mtg_now, mtg_new at row r+1
  double gam;         // relaxation coefficient at row r
};

template <int SCHEME, int LJ>
__global__ void __launch_bounds__(128, 4) stage_mv_kernel(const StageArgs a) {
  using F = Flux<SCHEME>;
  constexpr int E = F::extent;
  constexpr int NW = 2 * E;
  const int lane = threadIdx.x & 31;
  const int xw = (blockIdx.x + a.bx0) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (xw * MV_COLS >= a.nx) return;  // warp-uniform
  const int c = xw * MV_COLS + lane - 1;
  const int j0 = (blockIdx.y + a.by0) * LJ;
  const int jend = min(j0 + LJ, a.ny);
  const int k = blockIdx.z;
  const int nx = a.nx, ny = a.ny, nb = a.nb;

  const bool out_lane = lane >= 1 && lane <= MV_COLS && c < nx;
  const bool col_int = c >= nb && c < nx - nb;
  // clamped columns: cc keeps every x-offset load inside the row, cm is the lane's own
  // column whenever that exists (halo lanes beyond the domain only produce discarded values)
  const int cc = min(max(c, E), nx - E);
  const int cm = min(max(c, 0), nx - 1);

  const unsigned row = (unsigned)a.s_now.s1 * 8u;         // bytes per row (all 3-D fields)
  const unsigned plane = (unsigned)k * (unsigned)a.s_now.s2 * 8u;
  const unsigned grow = (unsigned)a.gamma.s1 * 8u;

  const double r_damp = a.damp ? a.rmat.ld(0, 0, k) : 0.0;
  const double one_m_eps = 1.0 - a.eps;

  const int r0 = j0 > 0 ? j0 - 1 : 0;  // first row computed (warm-up row unless j0 == 0)

  // ---- prologue: windows for the y-face r0 (rows r0-E .. r0+E-1) and its flux
  double wsu[NW], wsv[NW];
#pragma unroll
  for (int m = 0; m < NW; ++m) {
    const unsigned o = plane + (unsigned)max(r0 - E + m, 0) * row + (unsigned)cc * 8u;
    wsu[m] = ldo(a.su_int.p, o);
    wsv[m] = ldo(a.sv_int.p, o);
  }
  // running byte offsets of row r (advanced by `row` per iteration)
  unsigned o_cc = plane + (unsigned)r0 * row + (unsigned)cc * 8u;  // clamped column
  unsigned o_cm = plane + (unsigned)r0 * row + (unsigned)cm * 8u;  // own column
  unsigned o_g = (unsigned)r0 * grow + (unsigned)cm * 8u;          // gamma (2-D)
  double fy_su, fy_sv;
  {
    const double vq = F::prep(ldo(a.v_int.p, o_cc), a.fc);
    fy_su = F::eval_v(vq, wsu);
    fy_sv = F::eval_v(vq, wsv);
  }
  // Montgomery window: row r-1 (row r+1 arrives with the row loads)
  const unsigned o_m = plane + (unsigned)max(r0 - 1, 0) * row + (unsigned)cc * 8u;
  double mn_m = ldo(a.mtg_now.p, o_m), mn_0 = ldo(a.mtg_now.p, o_cc);
  double mw_m = ldo(a.mtg.p, o_m), mw_0 = ldo(a.mtg.p, o_cc);
  double sv_prev = 0.0, s_prev = 0.0;

  // rows up to jend + E are touched: inside the allocation because every field has at
  // least one more plane than nz (checked on the host); the values only reach discarded
  // boundary points
  auto load_row = [&](unsigned occ, unsigned ocm, unsigned og) {
    RowLoads L;
    L.su_w = ldo(a.su_int.p, occ + E * row);
    L.sv_w = ldo(a.sv_int.p, occ + E * row);
    L.v_n = ldo(a.v_int.p, occ + row);
    L.u_c = ldo(a.u_int.p, occ);
    L.s_pre = ldo(a.spre.p, ocm);
    L.s_now = ldo(a.s_now.p, ocm);
    L.su_now = ldo(a.su_now.p, ocm);
    L.sv_now = ldo(a.sv_now.p, ocm);
    L.mn_p = ldo(a.mtg_now.p, occ + row);
    L.mw_p = ldo(a.mtg.p, occ + row);
    L.gam = ldo(a.gamma.p, og);
    return L;
  };

  // DRAM -> L2 two rows further ahead, so the register loads above see L2 latency
  auto prefetch_row = [&](unsigned occ, unsigned ocm) {
    prefetch_l2(a.su_int.p, occ + E * row);
    prefetch_l2(a.sv_int.p, occ + E * row);
    prefetch_l2(a.v_int.p, occ + row);
    prefetch_l2(a.u_int.p, occ);
    prefetch_l2(a.spre.p, ocm);
    prefetch_l2(a.s_now.p, ocm);
    prefetch_l2(a.su_now.p, ocm);
    prefetch_l2(a.sv_now.p, ocm);
    prefetch_l2(a.mtg_now.p, occ + row);
    prefetch_l2(a.mtg.p, occ + row);
  };

  RowLoads nxt = load_row(o_cc, o_cm, o_g);
  for (int r = r0; r < jend; ++r) {
    const RowLoads cur = nxt;
    nxt = load_row(o_cc + row, o_cm + row, o_g + grow);  // in flight while row r is computed
    if (r + 3 < jend) prefetch_row(o_cc + 3 * row, o_cm + 3 * row);

    // ---- y-face r+1: shift the windows by one row, append row r+E
#pragma unroll
    for (int m = 0; m < NW - 1; ++m) {
      wsu[m] = wsu[m + 1];
      wsv[m] = wsv[m + 1];
    }
    wsu[NW - 1] = cur.su_w;
    wsv[NW - 1] = cur.sv_w;
    const double vq = F::prep(cur.v_n, a.fc);
    const double fy_su_p = F::eval_v(vq, wsu);
    const double fy_sv_p = F::eval_v(vq, wsv);

    // ---- left x-face of column c at row r: phi[c-E .. c+E-1]; phi[c] is wsu[E-1]
    // (the neighbours' lines entered L1 E rows ago as window loads)
    const double uq = F::prep(cur.u_c, a.fc);
    double xs[NW], ys[NW];
    {
      const double *psu = ptr_at(a.su_int.p, o_cc), *psv = ptr_at(a.sv_int.p, o_cc);
#pragma unroll
      for (int m = 0; m < NW; ++m) {
        xs[m] = m == E ? wsu[E - 1] : __ldg(psu + (m - E));
        ys[m] = m == E ? wsv[E - 1] : __ldg(psv + (m - E));
      }
    }
    const double fx_su = F::eval_v(uq, xs);
    const double fx_sv = F::eval_v(uq, ys);
    const double fx_su_p = __shfl_down_sync(0xffffffffu, fx_su, 1);
    const double fx_sv_p = __shfl_down_sync(0xffffffffu, fx_sv, 1);

    // ---- point update (prognostics/utils.py:L191-L204)
    const bool interior = col_int && r >= nb && r < ny - nb;
    double s = cur.s_pre, su = 0.0, sv = 0.0;
    if (interior) {
      {
        const double div = (fx_su_p - fx_su) / a.fc.dx + (fy_su_p - fy_su) / a.fc.dy;
        const double *pmn = ptr_at(a.mtg_now.p, o_cc), *pmw = ptr_at(a.mtg.p, o_cc);
        const double pg_now = one_m_eps * cur.s_now * (__ldg(pmn + 1) - __ldg(pmn - 1)) / a.two_dx;
        const double pg_new = a.eps * s * (__ldg(pmw + 1) - __ldg(pmw - 1)) / a.two_dx;
        su = cur.su_now - a.dt * (div + pg_now + pg_new - 0.0);
      }
      {
        const double div = (fx_sv_p - fx_sv) / a.fc.dx + (fy_sv_p - fy_sv) / a.fc.dy;
        const double pg_now = one_m_eps * cur.s_now * (cur.mn_p - mn_m) / a.two_dy;
        const double pg_new = a.eps * s * (cur.mw_p - mw_m) / a.two_dy;
        sv = cur.sv_now - a.dt * (div + pg_now + pg_new - 0.0);
      }
    }
    const double gam = cur.gam;
    double s_ref = 0.0, su_ref = 0.0, sv_ref = 0.0;
    if (gam != 0.0 || r_damp != 0.0) {
      s_ref = ldo(a.s_ref.p, o_cm);
      su_ref = ldo(a.su_ref.p, o_cm);
      sv_ref = ldo(a.sv_ref.p, o_cm);
    }
    if (!interior && gam != 1.0) {  // not reached with a Relaxed boundary (gamma == 1 there)
      su = ldo(a.su_new.p, o_cm);
      sv = ldo(a.sv_new.p, o_cm);
    }
    if (gam != 0.0) {  // hb.enforce_raw, dycore.py:L686
      s = relax_point(gam, s, s_ref);
      su = relax_point(gam, su, su_ref);
      sv = relax_point(gam, sv, sv_ref);
    }
    if (r_damp != 0.0) {  // dycore.py:L694-L700
      s = damp_point(cur.s_now, s, s_ref, r_damp, a.dt_full);
      su = damp_point(cur.su_now, su, su_ref, r_damp, a.dt_full);
      sv = damp_point(cur.sv_now, sv, sv_ref, r_damp, a.dt_full);
    }

    // ---- velocity diagnosis (dwarfs/diagnostics.py:L219-L272) and stores
    const double su_l = __shfl_up_sync(0xffffffffu, su, 1);
    const double s_l = __shfl_up_sync(0xffffffffu, s, 1);
    if (out_lane && r >= j0) {
      sto(a.s_new.p, o_cm, s);
      sto(a.su_new.p, o_cm, su);
      sto(a.sv_new.p, o_cm, sv);
      sto(a.u_new.p, o_cm, c == 0 ? ldo(a.u_ref.p, o_cm) : (su_l + su) / (s_l + s));
      if (c == nx - 1) sto(a.u_new.p, o_cm + 8u, ldo(a.u_ref.p, o_cm + 8u));  // relaxed.py:L161-L175
      sto(a.v_new.p, o_cm, r == 0 ? ldo(a.v_ref.p, o_cm) : (sv_prev + sv) / (s_prev + s));
      if (r == ny - 1) sto(a.v_new.p, o_cm + row, ldo(a.v_ref.p, o_cm + row));  // relaxed.py:L177-L191
    }
    sv_prev = sv;
    s_prev = s;
    fy_su = fy_su_p;
    fy_sv = fy_sv_p;
    mn_m = mn_0; mn_0 = cur.mn_p;
    mw_m = mw_0; mw_0 = cur.mw_p;
    o_cc += row; o_cm += row; o_g += grow;
  }
}

// TB200_STAGE_IMPL=tma selects the TMA / shared-memory-ring momentum kernel of
// isentropic_tma.cu instead of the register-window kernel below.  Both give bit-identical
// results; on B200 the register-window kernel is currently the faster one (1.94 ms vs 2.16 ms
// per launch at 1024x1024x64), hence the default.
int stage_impl() {
  static int impl = -1;
  if (impl < 0) {
    const char *e = getenv("TB200_STAGE_IMPL");
    impl = (e != nullptr && strcmp(e, "tma") == 0) ? 1 : 0;
  }
  return impl;
}

// TB200_S_IMPL=column selects the thread-per-column kernel S instead of kernels A + B
int s_impl() {
  static int impl = -1;
  if (impl < 0) {
    const char *e = getenv("TB200_S_IMPL");
    impl = (e != nullptr && strcmp(e, "column") == 0) ? 0 : 1;
  }
  return impl;
}

// optional per-kernel timing of the last stage call (tb200_stage_profile): four events on the
// launching stream around the (up to) three kernels
struct StageProfile {
  bool on = false, recorded = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
} g_prof;

void prof_mark(int n, cudaStream_t st) {
  if (g_prof.on) cudaEventRecord(g_prof.ev[n], st);
}

// ---------------------------------------------------------------- kernel MV, ring version
// Same decomposition and arithmetic as stage_mv_kernel (one warp = 30 columns x LJ rows of one
// level, marching in j), but the rotating state lives in WARP-PRIVATE shared-memory rings
// instead of register windows: each row of su_int / sv_int (and mtg_now / mtg_new) is loaded
// once with coalesced LDGs, stored into the warp's ring (lanes 0..4 add the five halo columns)
// and every stencil neighbour -- the six rows of the y stencil, the x neighbours, the Montgomery
// cross -- is then an LDS at an immediate offset from ONE per-row base.  The rings are mirrored
// (a row stored at slot s < W-1 is also stored at s + R) so that W consecutive rows are always
// contiguous: no modulo per access.  This removes the ~100 register moves and most of the
// address arithmetic per row of the register-window kernel; warps stay autonomous (one
// __syncwarp per row, no block barrier).
constexpr int RW = 38;                    // ring row: columns c_0-3 .. c_0+33 (+1 pad)
constexpr int SU_RING = 8, SU_MIRROR = 5; // y stencil spans up to 6 rows
constexpr int MT_RING = 4, MT_MIRROR = 2; // Montgomery rows r-1, r, r+1
constexpr int SU_ROWS = SU_RING + SU_MIRROR, MT_ROWS = MT_RING + MT_MIRROR;
constexpr int WARP_DOUBLES = (2 * SU_ROWS + 2 * MT_ROWS) * RW;

struct RingLoads {   // values on their way into the rings (row r+E of su/sv, row r+1 of mtg)
  double su, sv, mn, mw;      // own column
  double hsu, hsv, hmn, hmw;  // halo column of lanes 0..4
};
struct OwnLoads {    // own-column values without reuse, requested one row ahead
  double v_n, u_c, s_pre, s_now, su_now, sv_now, gam;
};

template <int SCHEME, int LJ>
__global__ void __launch_bounds__(128, 4) stage_mv_ring_kernel(const StageArgs a) {
  using F = Flux<SCHEME>;
  constexpr int E = F::extent;
  constexpr int NW = 2 * E;
  extern __shared__ double ring_smem[];
  const int lane = threadIdx.x & 31;
  const int xw = (blockIdx.x + a.bx0) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (xw * MV_COLS >= a.nx) return;  // warp-uniform (no block-wide barrier in this kernel)
  double *r_su = ring_smem + (threadIdx.x >> 5) * WARP_DOUBLES;
  double *r_sv = r_su + SU_ROWS * RW;
  double *r_mn = r_sv + SU_ROWS * RW;
  double *r_mw = r_mn + MT_ROWS * RW;
  const int c = xw * MV_COLS + lane - 1;
  const int j0 = (blockIdx.y + a.by0) * LJ;
  const int jend = min(j0 + LJ, a.ny);
  const int k = blockIdx.z;
  const int nx = a.nx, ny = a.ny, nb = a.nb;

  const bool out_lane = lane >= 1 && lane <= MV_COLS && c < nx;
  const bool col_int = c >= nb && c < nx - nb;
  const int cm = min(max(c, 0), nx - 1);  // own column, clamped into the row
  // halo column of lanes 0..4: ring columns 0, 1, 2 (left of lane 0) and 35, 36 (right of lane 31)
  const bool halo_lane = lane < 5;
  const int th = lane < 3 ? lane : 32 + lane;                       // ring column of the halo value
  const int ch = min(max(xw * MV_COLS - 1 - 3 + th, 0), nx - 1);     // its grid column, clamped
  const int t = lane + 3;                                            // ring column of the own value

  const unsigned row = (unsigned)a.s_now.s1 * 8u;
  const unsigned plane = (unsigned)k * (unsigned)a.s_now.s2 * 8u;
  const unsigned grow = (unsigned)a.gamma.s1 * 8u;
  const double r_damp = a.damp ? a.rmat.ld(0, 0, k) : 0.0;
  const double one_m_eps = 1.0 - a.eps;
  const int r0 = j0 > 0 ? j0 - 1 : 0;  // first row computed (warm-up row unless j0 == 0)

  // ring slots: su / sv row rho sits at slot (rho - (r0 - E)) & 7, mtg row rho at (rho - (r0 - 1)) & 3
  auto put_su = [&](int q, double vsu, double vsv, double hsu, double hsv) {
    const int s = q & (SU_RING - 1);
    r_su[s * RW + t] = vsu;
    r_sv[s * RW + t] = vsv;
    if (halo_lane) {
      r_su[s * RW + th] = hsu;
      r_sv[s * RW + th] = hsv;
    }
    if (s < SU_MIRROR) {
      r_su[(s + SU_RING) * RW + t] = vsu;
      r_sv[(s + SU_RING) * RW + t] = vsv;
      if (halo_lane) {
        r_su[(s + SU_RING) * RW + th] = hsu;
        r_sv[(s + SU_RING) * RW + th] = hsv;
      }
    }
  };
  auto put_mt = [&](int q, double vmn, double vmw, double hmn, double hmw) {
    const int s = q & (MT_RING - 1);
    r_mn[s * RW + t] = vmn;
    r_mw[s * RW + t] = vmw;
    if (halo_lane) {
      r_mn[s * RW + th] = hmn;
      r_mw[s * RW + th] = hmw;
    }
    if (s < MT_MIRROR) {
      r_mn[(s + MT_RING) * RW + t] = vmn;
      r_mw[(s + MT_RING) * RW + t] = vmw;
      if (halo_lane) {
        r_mn[(s + MT_RING) * RW + th] = hmn;
        r_mw[(s + MT_RING) * RW + th] = hmw;
      }
    }
  };

  // ---- prologue: su / sv rows r0-E .. r0+E-1 (ring rows 0 .. 2E-1), mtg rows r0-1, r0
  const unsigned col_o = (unsigned)cm * 8u, col_h = (unsigned)ch * 8u;
#pragma unroll
  for (int m = 0; m < NW; ++m) {
    const unsigned o = plane + (unsigned)max(r0 - E + m, 0) * row;
    put_su(m, ldo(a.su_int.p, o + col_o), ldo(a.sv_int.p, o + col_o),
           halo_lane ? ldo(a.su_int.p, o + col_h) : 0.0, halo_lane ? ldo(a.sv_int.p, o + col_h) : 0.0);
  }
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const unsigned o = plane + (unsigned)max(r0 - 1 + m, 0) * row;
    put_mt(m, ldo(a.mtg_now.p, o + col_o), ldo(a.mtg.p, o + col_o),
           halo_lane ? ldo(a.mtg_now.p, o + col_h) : 0.0, halo_lane ? ldo(a.mtg.p, o + col_h) : 0.0);
  }
  unsigned o_c = plane + (unsigned)r0 * row + col_o;  // own column, row r
  unsigned o_h = plane + (unsigned)r0 * row + col_h;  // halo column, row r
  unsigned o_g = (unsigned)r0 * grow + col_o;         // gamma (2-D)
  __syncwarp();
  double fy_su, fy_sv;
  {
    double ysu[NW], ysv[NW];
#pragma unroll
    for (int m = 0; m < NW; ++m) {
      ysu[m] = r_su[m * RW + t];
      ysv[m] = r_sv[m * RW + t];
    }
    const double vq = F::prep(ldo(a.v_int.p, o_c), a.fc);
    fy_su = F::eval_v(vq, ysu);
    fy_sv = F::eval_v(vq, ysv);
  }
  double sv_prev = 0.0, s_prev = 0.0;

  // rows up to jend + E are touched: inside the allocation (>= nz + 1 planes, checked on the host)
  auto load_ring = [&](unsigned oc, unsigned oh) {
    RingLoads L;
    L.su = ldo(a.su_int.p, oc + E * row);
    L.sv = ldo(a.sv_int.p, oc + E * row);
    L.mn = ldo(a.mtg_now.p, oc + row);
    L.mw = ldo(a.mtg.p, oc + row);
    L.hsu = L.hsv = L.hmn = L.hmw = 0.0;
    if (halo_lane) {
      L.hsu = ldo(a.su_int.p, oh + E * row);
      L.hsv = ldo(a.sv_int.p, oh + E * row);
      L.hmn = ldo(a.mtg_now.p, oh + row);
      L.hmw = ldo(a.mtg.p, oh + row);
    }
    return L;
  };
  auto load_own = [&](unsigned oc, unsigned og) {
    OwnLoads L;
    L.v_n = ldo(a.v_int.p, oc + row);
    L.u_c = ldo(a.u_int.p, oc);
    L.s_pre = ldo(a.spre.p, oc);
    L.s_now = ldo(a.s_now.p, oc);
    L.su_now = ldo(a.su_now.p, oc);
    L.sv_now = ldo(a.sv_now.p, oc);
    L.gam = ldo(a.gamma.p, og);
    return L;
  };
  auto prefetch_row = [&](unsigned oc) {
    prefetch_l2(a.su_int.p, oc + E * row);
    prefetch_l2(a.sv_int.p, oc + E * row);
    prefetch_l2(a.v_int.p, oc + row);
    prefetch_l2(a.u_int.p, oc);
    prefetch_l2(a.spre.p, oc);
    prefetch_l2(a.s_now.p, oc);
    prefetch_l2(a.su_now.p, oc);
    prefetch_l2(a.sv_now.p, oc);
    prefetch_l2(a.mtg_now.p, oc + row);
    prefetch_l2(a.mtg.p, oc + row);
  };

  RingLoads pend = load_ring(o_c, o_h);  // rows r0+E / r0+1: stored at the top of iteration r0
  OwnLoads nxt = load_own(o_c, o_g);
  for (int r = r0; r < jend; ++r) {
    const int q = r - r0;
    // ---- the rows requested during the previous iteration enter the rings ...
    put_su(q + NW, pend.su, pend.sv, pend.hsu, pend.hsv);
    put_mt(q + 2, pend.mn, pend.mw, pend.hmn, pend.hmw);
    // ... and the next ones are requested (in flight while row r is computed)
    pend = load_ring(o_c + row, o_h + row);
    const OwnLoads cur = nxt;
    nxt = load_own(o_c + row, o_g + grow);
    if (r + 3 < jend) prefetch_row(o_c + 3 * row);
    __syncwarp();

    // ---- one base per ring and row: every neighbour below is an immediate offset from it
    const double *bsu = r_su + ((q + 1) & (SU_RING - 1)) * RW + t;  // rows r-E+1 .. r+E
    const double *bsv = bsu + SU_ROWS * RW;
    const double *bmn = r_mn + (q & (MT_RING - 1)) * RW + t;        // rows r-1, r, r+1
    const double *bmw = bmn + MT_ROWS * RW;

    // ---- y-face r+1
    double ysu[NW], ysv[NW];
#pragma unroll
    for (int m = 0; m < NW; ++m) {
      ysu[m] = bsu[m * RW];
      ysv[m] = bsv[m * RW];
    }
    const double vq = F::prep(cur.v_n, a.fc);
    const double fy_su_p = F::eval_v(vq, ysu);
    const double fy_sv_p = F::eval_v(vq, ysv);

    // ---- left x-face of column c at row r (window entry E-1): phi[c-E .. c+E-1]
    const double uq = F::prep(cur.u_c, a.fc);
    double xs[NW], ys[NW];
#pragma unroll
    for (int m = 0; m < NW; ++m) {
      xs[m] = m == E ? ysu[E - 1] : bsu[(E - 1) * RW + (m - E)];
      ys[m] = m == E ? ysv[E - 1] : bsv[(E - 1) * RW + (m - E)];
    }
    const double fx_su = F::eval_v(uq, xs);
    const double fx_sv = F::eval_v(uq, ys);
    const double fx_su_p = __shfl_down_sync(0xffffffffu, fx_su, 1);
    const double fx_sv_p = __shfl_down_sync(0xffffffffu, fx_sv, 1);

    // ---- point update (prognostics/utils.py:L191-L204)
    const bool interior = col_int && r >= nb && r < ny - nb;
    double s = cur.s_pre, su = 0.0, sv = 0.0;
    if (interior) {
      {
        const double div = (fx_su_p - fx_su) / a.fc.dx + (fy_su_p - fy_su) / a.fc.dy;
        const double pg_now = one_m_eps * cur.s_now * (bmn[RW + 1] - bmn[RW - 1]) / a.two_dx;
        const double pg_new = a.eps * s * (bmw[RW + 1] - bmw[RW - 1]) / a.two_dx;
        su = cur.su_now - a.dt * (div + pg_now + pg_new - 0.0);
      }
      {
        const double div = (fx_sv_p - fx_sv) / a.fc.dx + (fy_sv_p - fy_sv) / a.fc.dy;
        const double pg_now = one_m_eps * cur.s_now * (bmn[2 * RW] - bmn[0]) / a.two_dy;
        const double pg_new = a.eps * s * (bmw[2 * RW] - bmw[0]) / a.two_dy;
        sv = cur.sv_now - a.dt * (div + pg_now + pg_new - 0.0);
      }
    }
    const double gam = cur.gam;
    double s_ref = 0.0, su_ref = 0.0, sv_ref = 0.0;
    if (gam != 0.0 || r_damp != 0.0) {
      s_ref = ldo(a.s_ref.p, o_c);
      su_ref = ldo(a.su_ref.p, o_c);
      sv_ref = ldo(a.sv_ref.p, o_c);
    }
    if (!interior && gam != 1.0) {  // not reached with a Relaxed boundary (gamma == 1 there)
      su = ldo(a.su_new.p, o_c);
      sv = ldo(a.sv_new.p, o_c);
    }
    if (gam != 0.0) {  // hb.enforce_raw, dycore.py:L686
      s = relax_point(gam, s, s_ref);
      su = relax_point(gam, su, su_ref);
      sv = relax_point(gam, sv, sv_ref);
    }
    if (r_damp != 0.0) {  // dycore.py:L694-L700
      s = damp_point(cur.s_now, s, s_ref, r_damp, a.dt_full);
      su = damp_point(cur.su_now, su, su_ref, r_damp, a.dt_full);
      sv = damp_point(cur.sv_now, sv, sv_ref, r_damp, a.dt_full);
    }

    // ---- velocity diagnosis (dwarfs/diagnostics.py:L219-L272) and stores
    const double su_l = __shfl_up_sync(0xffffffffu, su, 1);
    const double s_l = __shfl_up_sync(0xffffffffu, s, 1);
    if (out_lane && r >= j0) {
      sto(a.s_new.p, o_c, s);
      sto(a.su_new.p, o_c, su);
      sto(a.sv_new.p, o_c, sv);
      sto(a.u_new.p, o_c, c == 0 ? ldo(a.u_ref.p, o_c) : (su_l + su) / (s_l + s));
      if (c == nx - 1) sto(a.u_new.p, o_c + 8u, ldo(a.u_ref.p, o_c + 8u));  // relaxed.py:L161-L175
      sto(a.v_new.p, o_c, r == 0 ? ldo(a.v_ref.p, o_c) : (sv_prev + sv) / (s_prev + s));
      if (r == ny - 1) sto(a.v_new.p, o_c + row, ldo(a.v_ref.p, o_c + row));  // relaxed.py:L177-L191
    }
    sv_prev = sv;
    s_prev = s;
    fy_su = fy_su_p;
    fy_sv = fy_sv_p;
    o_c += row; o_h += row; o_g += grow;
  }
}

// TB200_MV_IMPL=window selects the register-window momentum kernel instead of the ring one
int mv_impl() {
  static int impl = -1;
  if (impl < 0) {
    const char *e = getenv("TB200_MV_IMPL");
    impl = (e != nullptr && strcmp(e, "window") == 0) ? 0 : 1;
  }
  return impl;
}

// The momentum kernel over a rectangle of its block grid.
template <int SCHEME, int LJ, int WARPS>
int launch_mv_rect(StageArgs a, int bx0, int bx1, int by0, int by1, cudaStream_t st) {
  if (bx1 <= bx0 || by1 <= by0) return TB200_OK;
  a.bx0 = bx0;
  a.by0 = by0;
  dim3 block(32 * WARPS, 1, 1);
  dim3 grid(bx1 - bx0, by1 - by0, a.nz);
  if (mv_impl() != 0) {
    stage_mv_ring_kernel<SCHEME, LJ><<<grid, block, WARPS * WARP_DOUBLES * sizeof(double), st>>>(a);
    return check_launch("isentropic_stage_dry/MV(ring)");
  }
  stage_mv_kernel<SCHEME, LJ><<<grid, block, 0, st>>>(a);
  return check_launch("isentropic_stage_dry/MV");
}

// part 0: the whole block grid; part 1: the blocks holding the a.rim columns / rows next to an
// edge with a neighbour (south and north strips over all columns, west and east blocks over the
// remaining strips); part 2: the interior rectangle.
template <int SCHEME, int LJ, int WARPS>
int launch_mv_part(const StageArgs &a, int gx, int gy, cudaStream_t st) {
  if (a.part == 0) return launch_mv_rect<SCHEME, LJ, WARPS>(a, 0, gx, 0, gy, st);
  const int cols = WARPS * MV_COLS;
  // first interior block after the west rim / first east-rim block, likewise for the strips
  int xw = a.rim[0] > 0 ? (a.rim[0] + cols - 1) / cols : 0;
  int xe = a.rim[1] > 0 ? (a.nx - a.rim[1]) / cols : gx;
  int ys = a.rim[2] > 0 ? (a.rim[2] + LJ - 1) / LJ : 0;
  int yn = a.rim[3] > 0 ? (a.ny - a.rim[3]) / LJ : gy;
  xe = max(min(xe, gx), xw);
  yn = max(min(yn, gy), ys);
  if (a.part == 2) return launch_mv_rect<SCHEME, LJ, WARPS>(a, xw, xe, ys, yn, st);
  int rc = launch_mv_rect<SCHEME, LJ, WARPS>(a, 0, gx, 0, ys, st);
  if (!rc) rc = launch_mv_rect<SCHEME, LJ, WARPS>(a, 0, gx, yn, gy, st);
  if (!rc) rc = launch_mv_rect<SCHEME, LJ, WARPS>(a, 0, xw, ys, yn, st);
  if (!rc) rc = launch_mv_rect<SCHEME, LJ, WARPS>(a, xe, gx, ys, yn, st);
  return rc;
}

template <int SCHEME>
int run_stage(const StageArgs &a, cudaStream_t st) {
  if (a.part == 2) {  // the s-step and the scans ran with part 1
    constexpr int LJ = 64, WARPS = 4;
    const int chunks = (a.nx + MV_COLS - 1) / MV_COLS;
    return launch_mv_part<SCHEME, LJ, WARPS>(a, (chunks + WARPS - 1) / WARPS, (a.ny + LJ - 1) / LJ, st);
  }
  prof_mark(0, st);
  if (s_impl() != 0 && a.nz <= 64) {
    {
      constexpr int LJ = 64, WARPS = 4;
      const int chunks = (a.nx + A_COLS - 1) / A_COLS;
      dim3 block(32 * WARPS, 1, 1);
      dim3 grid((chunks + WARPS - 1) / WARPS, (a.ny + LJ - 1) / LJ, a.nz);
      stage_a_kernel<SCHEME, LJ><<<grid, block, 0, st>>>(a);
      int rc = check_launch("isentropic_stage_dry/A");
      if (rc) return rc;
      prof_mark(1, st);
    }
    {
      dim3 block(32, 4, 1);
      dim3 grid((a.nx + 31) / 32, (a.ny + 3) / 4, 1);
      if (a.nz == 64)
        stage_b_kernel<64, true><<<grid, block, 0, st>>>(a);
      else if (a.nz <= 32)
        stage_b_kernel<32, false><<<grid, block, 0, st>>>(a);
      else
        stage_b_kernel<64, false><<<grid, block, 0, st>>>(a);
      int rc = check_launch("isentropic_stage_dry/B");
      if (rc) return rc;
    }
  } else {
    dim3 block(32, 4, 1);
    dim3 grid((a.nx + 31) / 32, (a.ny + 3) / 4, 1);
    stage_s_kernel<SCHEME><<<grid, block, 0, st>>>(a);
    int rc = check_launch("isentropic_stage_dry/S");
    if (rc) return rc;
    prof_mark(1, st);
  }
  prof_mark(2, st);
  if (stage_impl() != 0 && a.part == 0) {  // TMA shared-memory-ring kernel; -1 = not covered
    const int rc = launch_stage_c(a, SCHEME, st);
    if (rc >= 0) {
      prof_mark(3, st);
      g_prof.recorded = g_prof.on;
      return rc;
    }
  }
  {
    constexpr int LJ = 64, WARPS = 4;
    const int chunks = (a.nx + MV_COLS - 1) / MV_COLS;
    const int gx = (chunks + WARPS - 1) / WARPS, gy = (a.ny + LJ - 1) / LJ;
    const int rc = launch_mv_part<SCHEME, LJ, WARPS>(a, gx, gy, st);
    prof_mark(3, st);
    g_prof.recorded = g_prof.on;
    return rc;
  }
}

bool covers(const View &v, int ni, int nj, int nk) {
  return v.ok() && v.n0 >= ni && v.n1 >= nj && v.n2 >= nk;
}

}  // namespace

extern "C" int tb200_stage_profile(int enable) {
  if (enable && g_prof.ev[0] == nullptr) {
    for (auto &e : g_prof.ev) {
      if (cudaEventCreate(&e) != cudaSuccess) {
        set_error("stage_profile: %s", cudaGetErrorString(cudaGetLastError()));
        return TB200_ERR_CUDA;
      }
    }
  }
  g_prof.on = enable != 0;
  g_prof.recorded = false;
  return TB200_OK;
}

extern "C" int tb200_stage_profile_read(double ms[3]) {
  TB200_REQUIRE(ms != nullptr && g_prof.recorded, "stage_profile_read: no profiled stage call yet");
  if (cudaEventSynchronize(g_prof.ev[3]) != cudaSuccess) {
    set_error("stage_profile_read: %s", cudaGetErrorString(cudaGetLastError()));
    return TB200_ERR_CUDA;
  }
  for (int n = 0; n < 3; ++n) {
    float t = 0.f;
    cudaEventElapsedTime(&t, g_prof.ev[n], g_prof.ev[n + 1]);
    ms[n] = t;
  }
  return TB200_OK;
}

extern "C" int tb200_isentropic_stage_dry(
    const tb200_isentropic_stage *cfg, const tb200_field *s_now, const tb200_field *su_now,
    const tb200_field *sv_now, const tb200_field *mtg_now, const tb200_field *s_int,
    const tb200_field *su_int, const tb200_field *sv_int, const tb200_field *u_int,
    const tb200_field *v_int, tb200_field *s_new, tb200_field *su_new, tb200_field *sv_new,
    tb200_field *u_new, tb200_field *v_new, const tb200_field *s_ref, const tb200_field *su_ref,
    const tb200_field *sv_ref, const tb200_field *u_ref, const tb200_field *v_ref,
    const tb200_field *gamma, const tb200_field *rmat, const tb200_field *hs,
    tb200_field *scratch_exn, tb200_field *scratch_mtg, tb200_field *scratch_s, void *stream) {
  TB200_REQUIRE(cfg != nullptr, "isentropic_stage_dry: NULL cfg");
  StageArgs a{};
  a.s_now = view(s_now); a.su_now = view(su_now); a.sv_now = view(sv_now);
  a.mtg_now = view(mtg_now);
  a.s_int = view(s_int); a.su_int = view(su_int); a.sv_int = view(sv_int);
  a.u_int = view(u_int); a.v_int = view(v_int);
  a.s_new = view(s_new); a.su_new = view(su_new); a.sv_new = view(sv_new);
  a.u_new = view(u_new); a.v_new = view(v_new);
  a.s_ref = view(s_ref); a.su_ref = view(su_ref); a.sv_ref = view(sv_ref);
  a.u_ref = view(u_ref); a.v_ref = view(v_ref);
  a.gamma = view(gamma); a.rmat = view(rmat); a.hs = view(hs);
  a.exn = view(scratch_exn); a.mtg = view(scratch_mtg); a.spre = view(scratch_s);
  a.nx = cfg->nx; a.ny = cfg->ny; a.nz = cfg->nz; a.nb = cfg->nb; a.damp = cfg->damp;
  a.part = cfg->part;
  for (int n = 0; n < 4; ++n) a.rim[n] = cfg->rim[n];
  TB200_REQUIRE(a.part >= 0 && a.part <= 2, "isentropic_stage_dry: part must be 0, 1 or 2");
  a.dt = cfg->dt; a.dt_full = cfg->dt_full; a.dx = cfg->dx; a.dy = cfg->dy; a.dz = cfg->dz;
  a.eps = cfg->eps; a.pt = cfg->pt; a.theta_s = cfg->theta_s;
  a.pref = cfg->constants[0]; a.rd = cfg->constants[1]; a.g = cfg->constants[2];
  a.cp = cfg->constants[3];
  a.fc = make_flux_const(a.dx, a.dy);
  a.two_dx = make_cdiv(2.0 * a.dx);
  a.two_dy = make_cdiv(2.0 * a.dy);
  a.cpref = make_cdiv(a.pref);

  const int nx = a.nx, ny = a.ny, nz = a.nz;
  int e = -1;
  switch (cfg->flux_scheme) {
    case TB200_FLUX_UPWIND: case TB200_FLUX_CENTERED: e = 1; break;
    case TB200_FLUX_THIRD_ORDER_UPWIND: e = 2; break;
    case TB200_FLUX_FIFTH_ORDER_UPWIND: e = 3; break;
  }
  TB200_REQUIRE(e > 0, "isentropic_stage_dry: unknown flux scheme %d", cfg->flux_scheme);
  TB200_REQUIRE(nz >= 1 && a.nb >= e && nx >= 2 * a.nb + 1 && ny >= 2 * a.nb + 1,
                "isentropic_stage_dry: need nb >= extent and nx, ny >= 2 nb + 1");
  const View *mass[] = {&a.s_now, &a.su_now, &a.sv_now, &a.mtg_now, &a.s_int, &a.su_int,
                        &a.sv_int, &a.s_new,  &a.su_new, &a.sv_new,  &a.s_ref, &a.su_ref,
                        &a.sv_ref, &a.exn,    &a.mtg,    &a.spre};
  for (const View *v : mass)
    TB200_REQUIRE(covers(*v, nx, ny, nz), "isentropic_stage_dry: a mass-point field is NULL or too small");
  {
    const View *all[] = {&a.s_now, &a.su_now, &a.sv_now, &a.mtg_now, &a.s_int, &a.su_int, &a.sv_int,
                         &a.u_int, &a.v_int, &a.s_new, &a.su_new, &a.sv_new, &a.u_new, &a.v_new,
                         &a.s_ref, &a.su_ref, &a.sv_ref, &a.u_ref, &a.v_ref, &a.exn, &a.mtg, &a.spre};
    for (const View *v : all) {
      if (v->ok() && v->s0 != 1) {
        set_error("isentropic_stage_dry: fields must have unit stride along i (b200 storage layout)");
        return TB200_ERR_LAYOUT;
      }
    }
  }
  TB200_REQUIRE(covers(a.u_int, nx + 1, ny, nz) && covers(a.u_new, nx + 1, ny, nz) &&
                    covers(a.u_ref, nx + 1, ny, nz),
                "isentropic_stage_dry: u fields must cover (nx+1, ny, nz)");
  TB200_REQUIRE(covers(a.v_int, nx, ny + 1, nz) && covers(a.v_new, nx, ny + 1, nz) &&
                    covers(a.v_ref, nx, ny + 1, nz),
                "isentropic_stage_dry: v fields must cover (nx, ny+1, nz)");
  TB200_REQUIRE(covers(a.gamma, nx, ny, 1) && covers(a.hs, nx, ny, 1),
                "isentropic_stage_dry: gamma / hs must cover (nx, ny, 1)");
  TB200_REQUIRE(!a.damp || covers(a.rmat, 1, 1, nz), "isentropic_stage_dry: rmat must cover (1, 1, nz)");
  TB200_REQUIRE(a.s_new.p != a.s_int.p && a.su_new.p != a.su_int.p && a.sv_new.p != a.sv_int.p &&
                    a.s_new.p != a.s_now.p && a.u_new.p != a.u_int.p && a.v_new.p != a.v_int.p,
                "isentropic_stage_dry: output fields must not alias the stage inputs");
  {
    const View *all[] = {&a.s_now, &a.su_now, &a.sv_now, &a.mtg_now, &a.s_int, &a.su_int, &a.sv_int,
                         &a.u_int, &a.v_int, &a.s_new, &a.su_new, &a.sv_new, &a.u_new, &a.v_new,
                         &a.s_ref, &a.su_ref, &a.sv_ref, &a.u_ref, &a.v_ref, &a.exn, &a.mtg, &a.spre};
    for (const View *v : all) {
      if (v->s1 != a.s_now.s1 || v->s2 != a.s_now.s2 || v->n2 < nz + 1 || v->n1 < ny + 1 ||
          v->s2 < v->s1 * (ny + 1) || (long long)v->s2 * v->n2 * 8 >= (1LL << 32)) {
        set_error("isentropic_stage_dry: all 3-D fields must share one geometry (equal row/plane "
                  "strides, >= nz+1 planes of >= ny+1 rows, < 4 GiB) -- allocate them with the "
                  "b200 allocator and one storage shape");
        return TB200_ERR_LAYOUT;
      }
    }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (cfg->flux_scheme) {
    case TB200_FLUX_UPWIND: return run_stage<TB200_FLUX_UPWIND>(a, st);
    case TB200_FLUX_CENTERED: return run_stage<TB200_FLUX_CENTERED>(a, st);
    case TB200_FLUX_THIRD_ORDER_UPWIND: return run_stage<TB200_FLUX_THIRD_ORDER_UPWIND>(a, st);
    default: return run_stage<TB200_FLUX_FIFTH_ORDER_UPWIND>(a, st);
  }
}
