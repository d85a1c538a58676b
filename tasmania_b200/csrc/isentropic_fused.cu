// isentropic_fused.cu -- one Runge-Kutta stage of the dry isentropic dynamical core with the
// relaxed lateral boundary, fused into three kernels (the benchmark hot path).
//
// Reference sequence per stage (src/tasmania/isentropic/dynamics/dycore.py:L641-L721 and
// subclasses/prognostics/rk3ws_si.py:L105-L234):
//   K1 step s -> irelax(s) -> montgomery(s_new) -> K2 step su, sv -> irelax(s, su, sv, u, v)
//   -> Rayleigh damping (s, su, sv) -> velocity_x / velocity_y -> outermost layers of u, v.
// The reference runs 13 full-domain passes for this; here three kernels (default path):
//
//   kernel A   (one warp per 31 columns x 64 rows of one level, marching in j)
//       s_pre = irelax(K1(...))                                -> scratch_s (= s_new: in place)
//   kernel B   (one thread per column, the column in registers)
//       pressure by the downward scan on s_pre, exn = cp (p/pref)^kappa (pow_pos_n<8>),
//       mtg_new by the upward scan                             -> scratch_mtg
//   kernel MV  (one warp per 60 columns x 64 rows of one level, two columns per lane, marching
//       in j; su / sv / mtg rows staged with cp.async into warp-private shared-memory rings)
//       su, sv = irelax(K2(...)), s = irelax(s_pre), Rayleigh damping on all three.
//
// The vertical scans are inherently two sweeps (pressure top-down, Montgomery bottom-up) and
// the momentum step needs mtg_new at i+-1 / j+-1, hence the kernel boundaries.
//
// Velocities (round 2).  The reference ends every stage with the diagnosis of u, v and starts the
// next one by reading them.  With derive_uv_in / skip_uv_out (tb200_isentropic_stage) no stage
// writes them: kernels A and MV re-diagnose the advecting velocities from s_int, su_int, sv_int
// with the same formula (same bits; the division is common.cuh:qdiv, which keeps zero numerators
// off the compiler's slow path), and the velocities of the step's final state come from one pass
// of tb200_velocity_components (elementwise.cu).  The moist stage (tb200_isentropic_stage_moist)
// adds kernel T between A and B: the three water constituents (density, K1 share, mass fraction,
// relaxation) in one tiled kernel.
//
// Periodic boundary (tb200_isentropic_stage.periodic).  The kernels are general in gamma: a point
// outside the interior keeps its previous value unless gamma == 1 there.  With gamma = 0 everywhere
// on the numerical grid of a periodic domain the same kernels therefore compute the interior and
// leave the ghost layers alone; the stage wraps s_pre into them between kernel A and kernel B (the
// reference's hb.enforce_field(s_new) before the Montgomery scan), and the caller wraps the outputs
// and applies the damping after the call, in the reference's order (dycore.py:L684-L700).
//
// HBM traffic per point and stage (8-byte words), stages 1, 2: A reads s_now, s_int, su_int,
// sv_int, writes s_pre (5); B reads s_pre, writes mtg (2); MV reads s_now, s_int, s_pre, mtg_now,
// mtg_new, su_now, su_int, sv_now, sv_int, writes su, sv and s where relaxation / damping change
// it (11+): 18 words = 144 B against the algorithmic minimum of 112 B (80 B
// without the u, v round trip); measured 164 B (146 B at stage 0).  All
// arithmetic follows the reference's operation order (see stencil_math.cuh).
//
// Earlier variants are kept selectable and bit-identical (tests/test_gpu_stage_variants.py):
// TB200_S_IMPL=column (kernel S = A + B in one thread-per-column kernel, pressures parked in
// memory), TB200_A_IMPL=two (two columns per lane), TB200_MV_IMPL=window | ring (register windows /
// one column per lane), TB200_MV_BLOCK=3x1 | 6x1, TB200_STAGE_IMPL=tma (isentropic_tma.cu),
// TB200_LAZY_UV=0 (the reference's velocity round trip), TB200_T_IMPL=point.
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

// DERIVE: the advecting velocities are re-diagnosed from s_int, su_int, sv_int (velocity_x /
// velocity_y, dwarfs/diagnostics.py:L219-L272) instead of read from u_int / v_int: the previous
// stage then need not write them (tb200_isentropic_stage.derive_uv_in).  Same traffic for this
// kernel (su_int, sv_int instead of u, v), two IEEE divisions more per point.
// TND: a slow tendency of s is passed (prognostics/utils.py:L95-L99: s_now - dt (div - s_tnd))
template <int SCHEME, int LJ, bool DERIVE, bool TND = false>
__global__ void __launch_bounds__(128, (DERIVE || TND) ? 6 : 7) stage_a_kernel(const StageArgs a) {
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
  // v at the y-face j is (sv[j-1] + sv[j]) / (s[j-1] + s[j]); sv_lo carries sv_int of the row below
  double sv_lo = 0.0;
  double fy;
  if (DERIVE) {
    sv_lo = ldo(a.sv_int.p, o_cc);
    const double sv_m = ldo(a.sv_int.p, plane + (unsigned)max(j0 - 1, 0) * row + (unsigned)cc * 8u);
    fy = F::eval_v(F::prep(qdiv(sv_m + sv_lo, ws[E - 1] + ws[E]), a.fc), ws);
  } else {
    fy = F::eval_v(F::prep(ldo(a.v_int.p, o_cc), a.fc), ws);
  }

  struct Row {
    double s_w, v_n, u_c, s_now, gam;  // DERIVE: v_n = sv_int at row r+1, u_c = su_int at row r
  };
  auto load_row = [&](unsigned occ, unsigned ocm, unsigned og) {
    Row L;
    L.s_w = ldo(a.s_int.p, occ + E * row);
    L.v_n = ldo(DERIVE ? a.sv_int.p : a.v_int.p, occ + row);
    L.u_c = ldo(DERIVE ? a.su_int.p : a.u_int.p, occ);
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
      prefetch_l2(DERIVE ? a.sv_int.p : a.v_int.p, o_cc + 5 * row);
      prefetch_l2(DERIVE ? a.su_int.p : a.u_int.p, o_cc + 4 * row);
      prefetch_l2(a.s_now.p, o_cm + 4 * row);
    }
#pragma unroll
    for (int m = 0; m < NW - 1; ++m) ws[m] = ws[m + 1];
    ws[NW - 1] = cur.s_w;  // ws[m] = s_int at row r - E + 1 + m
    double vq;
    if (DERIVE) {
      vq = F::prep(qdiv(sv_lo + cur.v_n, ws[E - 1] + ws[E]), a.fc);
      sv_lo = cur.v_n;
    } else {
      vq = F::prep(cur.v_n, a.fc);
    }
    const double fy_p = F::eval_v(vq, ws);
    // x-neighbours of s_int from L1 (requesting them a row ahead like the first-touch streams was
    // measured SLOWER: 0.54 / 0.80 against 0.51 / 0.73 ms, the ten extra live registers spill in
    // the DERIVE variant; profiles/README.md round 2)
    double xs[NW];
    {
      const double *ps = ptr_at(a.s_int.p, o_cc);
#pragma unroll
      for (int m = 0; m < NW; ++m) xs[m] = m == E ? ws[E - 1] : __ldg(ps + (m - E));
    }
    double uq;
    if (DERIVE) {  // u at the left face of the column: (su[c-1] + su[c]) / (s[c-1] + s[c])
      // (the left column from L1 like the s_int neighbours: a shuffle would hand over the
      // CLAMPED column of a lane next to the domain edge)
      const double su_l = ldo(a.su_int.p, o_cc - 8u);
      uq = F::prep(qdiv(su_l + cur.u_c, xs[E - 1] + xs[E]), a.fc);
    } else {
      uq = F::prep(cur.u_c, a.fc);
    }
    const double fx = F::eval_v(uq, xs);
    const double fx_p = __shfl_down_sync(0xffffffffu, fx, 1);

    const bool interior = col_int && r >= nb && r < ny - nb;
    const double gam = cur.gam;
    double v;
    if (interior) {  // prognostics/utils.py:L95-L99
      const double div = (fx_p - fx) / a.fc.dx + (fy_p - fy) / a.fc.dy;
      const double tnd = TND ? ldo(a.s_tnd.p, o_cm) : 0.0;
      v = cur.s_now - a.dt * (div - tnd);
    } else {
      v = gam == 1.0 ? 0.0 : ldo(a.s_new.p, o_cm);  // untouched by K1; the relaxation decides
    }
    if (gam != 0.0) v = relax_point(gam, v, ldo(a.s_ref.p, o_cm));  // rk3ws_si.py:L184-L189
    if (out_lane) sto(a.spre.p, o_cm, v);
    fy = fy_p;
    o_cc += row; o_cm += row; o_g += grow;
  }
}

// ---------------------------------------------------------------- kernel A, two columns per lane
// Same decomposition and arithmetic as stage_a_kernel (a warp marches along j over a strip of one
// level, register window of s_int rows, every face flux evaluated once), but a lane owns an ALIGNED
// PAIR of columns like the momentum kernel: every access is LDG.128 / STG.128, a warp has twice
// the bytes in flight per load, and address arithmetic, loop control and the relaxation logic are
// paid once per two points.  One warp = 64 columns, 62 of them owned (lane 31 lends the left face
// of its pair).  The x-neighbours of the pair are the pairs at c0 - 4, c0 - 2 and c0 + 2 (L1).
// Requirements as for the two-column momentum kernel (mv2_ok).  NOT the default
// (TB200_A_IMPL=two selects it): measured SLOWER than the one-column kernel at config 5 -- 0.575 /
// 0.87 ms against 0.51 / 0.735 ms per launch (round 2): 118-126 registers leave 16 warps per SM
// where the one-column kernel has 28, and this kernel lives on warps in flight, not on
// instruction count.  Kept as a tested variant (bit-identical, tests/test_gpu_stage_variants.py).
constexpr int A2_COLS = 62;
__device__ __forceinline__ double2 ld2o(const double *base, unsigned off) {
  return __ldg(reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(base) + off));
}

template <int SCHEME, int LJ, bool DERIVE>
__global__ void __launch_bounds__(128, 4) stage_a2_kernel(const StageArgs a) {
  using F = Flux<SCHEME>;
  constexpr int E = F::extent;
  constexpr int NW = 2 * E;
  const int lane = threadIdx.x & 31;
  const int xw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (xw * A2_COLS >= a.nx) return;  // warp-uniform
  const int c0 = xw * A2_COLS + 2 * lane;  // even
  const int j0 = blockIdx.y * LJ;
  const int jend = min(j0 + LJ, a.ny);
  const int k = blockIdx.z;
  const int nx = a.nx, ny = a.ny, nb = a.nb;
  const int pmax = (nx - 1) & ~1;  // last aligned pair that starts inside the row
  const bool own = lane < 31;
  const bool out0 = own && c0 < nx, out1 = own && c0 + 1 < nx;
  const bool int0 = c0 >= nb && c0 < nx - nb, int1 = c0 + 1 >= nb && c0 + 1 < nx - nb;
  // pairs: own, and the x-neighbours (clamped into the row: a clamped pair only feeds faces of
  // points outside the computational domain)
  const unsigned col = (unsigned)min(c0, pmax) * 8u;
  const unsigned col_m4 = (unsigned)min(max(c0 - 4, 0), pmax) * 8u;
  const unsigned col_m2 = (unsigned)min(max(c0 - 2, 0), pmax) * 8u;
  const unsigned col_p2 = (unsigned)min(c0 + 2, pmax) * 8u;
  const unsigned col_l1 = (unsigned)min(max(c0 - 1, 0), nx - 1) * 8u;  // DERIVE: su_int left of the pair

  const unsigned row = (unsigned)a.s_now.s1 * 8u;
  const unsigned plane = (unsigned)k * (unsigned)a.s_now.s2 * 8u;
  const unsigned grow = (unsigned)a.gamma.s1 * 8u;
  const double2 zero2 = make_double2(0.0, 0.0);

  // window for the y-face j0 (rows j0-E .. j0+E-1) and its flux
  double2 ws[NW];
#pragma unroll
  for (int m = 0; m < NW; ++m) ws[m] = ld2o(a.s_int.p, plane + (unsigned)max(j0 - E + m, 0) * row + col);
  unsigned o = plane + (unsigned)j0 * row;  // row r, column 0
  unsigned og = (unsigned)j0 * grow + col;
  double2 sv_lo = zero2;
  double fy0, fy1;
  {
    double w0[NW], w1[NW];
#pragma unroll
    for (int m = 0; m < NW; ++m) { w0[m] = ws[m].x; w1[m] = ws[m].y; }
    double vq0, vq1;
    if (DERIVE) {
      sv_lo = ld2o(a.sv_int.p, o + col);
      const double2 sv_m = ld2o(a.sv_int.p, plane + (unsigned)max(j0 - 1, 0) * row + col);
      vq0 = F::prep(qdiv(sv_m.x + sv_lo.x, ws[E - 1].x + ws[E].x), a.fc);
      vq1 = F::prep(qdiv(sv_m.y + sv_lo.y, ws[E - 1].y + ws[E].y), a.fc);
    } else {
      const double2 v0 = ld2o(a.v_int.p, o + col);
      vq0 = F::prep(v0.x, a.fc);
      vq1 = F::prep(v0.y, a.fc);
    }
    fy0 = F::eval_v(vq0, w0);
    fy1 = F::eval_v(vq1, w1);
  }

  struct Row {
    double2 s_w, v_n, u_c, s_now, gam;  // DERIVE: v_n = sv_int at row r+1, u_c = su_int at row r
  };
  auto load_row = [&](unsigned orow, unsigned ogam) {
    Row L;
    L.s_w = ld2o(a.s_int.p, orow + E * row + col);
    L.v_n = ld2o(DERIVE ? a.sv_int.p : a.v_int.p, orow + row + col);
    L.u_c = ld2o(DERIVE ? a.su_int.p : a.u_int.p, orow + col);
    L.s_now = ld2o(a.s_now.p, orow + col);
    L.gam = ld2o(a.gamma.p, ogam);
    return L;
  };
  Row nxt = load_row(o, og);
  for (int r = j0; r < jend; ++r) {
    const Row cur = nxt;
    nxt = load_row(o + row, og + grow);
    if (r + 4 < jend && (lane & 7) == 0) {  // DRAM -> L2 a few rows ahead, one request per 128-byte line
      prefetch_l2(a.s_int.p, o + (E + 4) * row + col);
      prefetch_l2(DERIVE ? a.sv_int.p : a.v_int.p, o + 5 * row + col);
      prefetch_l2(DERIVE ? a.su_int.p : a.u_int.p, o + 4 * row + col);
      prefetch_l2(a.s_now.p, o + 4 * row + col);
    }
#pragma unroll
    for (int m = 0; m < NW - 1; ++m) ws[m] = ws[m + 1];
    ws[NW - 1] = cur.s_w;  // ws[m] = s_int at row r - E + 1 + m
    // ---- y-faces r+1 of both columns
    double vq0, vq1;
    if (DERIVE) {
      vq0 = F::prep(qdiv(sv_lo.x + cur.v_n.x, ws[E - 1].x + ws[E].x), a.fc);
      vq1 = F::prep(qdiv(sv_lo.y + cur.v_n.y, ws[E - 1].y + ws[E].y), a.fc);
      sv_lo = cur.v_n;
    } else {
      vq0 = F::prep(cur.v_n.x, a.fc);
      vq1 = F::prep(cur.v_n.y, a.fc);
    }
    double fy0_p, fy1_p;
    {
      double w0[NW], w1[NW];
#pragma unroll
      for (int m = 0; m < NW; ++m) { w0[m] = ws[m].x; w1[m] = ws[m].y; }
      fy0_p = F::eval_v(vq0, w0);
      fy1_p = F::eval_v(vq1, w1);
    }
    // ---- x-faces at row r: xc[m] = s_int at column c0 - E + m, m = 0 .. 2E
    double xc[NW + 1];
    {
      const double2 ownp = ws[E - 1];
      const double2 n2 = ld2o(a.s_int.p, o + col_m2), p2 = ld2o(a.s_int.p, o + col_p2);
      double2 n4 = zero2;
      if (E >= 3) n4 = ld2o(a.s_int.p, o + col_m4);
#pragma unroll
      for (int m = 0; m <= NW; ++m) {
        const int d = m - E;
        xc[m] = d == -3 ? n4.y : d == -2 ? n2.x : d == -1 ? n2.y : d == 0 ? ownp.x : d == 1 ? ownp.y
                : d == 2 ? p2.x : p2.y;
      }
    }
    double uq0, uq1;
    if (DERIVE) {  // u at the left face of c0 and at the face between c0 and c1
      const double su_l = ldo(a.su_int.p, o + col_l1);
      uq0 = F::prep(qdiv(su_l + cur.u_c.x, xc[E - 1] + xc[E]), a.fc);
      uq1 = F::prep(qdiv(cur.u_c.x + cur.u_c.y, xc[E] + xc[E + 1]), a.fc);
    } else {
      uq0 = F::prep(cur.u_c.x, a.fc);
      uq1 = F::prep(cur.u_c.y, a.fc);
    }
    const double fx0 = F::eval_v(uq0, xc), fx1 = F::eval_v(uq1, xc + 1);
    const double fx2 = __shfl_down_sync(0xffffffffu, fx0, 1);

    // ---- point updates (prognostics/utils.py:L95-L99) and first relaxation (rk3ws_si.py:L184-L189)
    const bool row_int = r >= nb && r < ny - nb;
    double v0, v1;
    if (row_int && int0) {
      const double div = (fx1 - fx0) / a.fc.dx + (fy0_p - fy0) / a.fc.dy;
      v0 = cur.s_now.x - a.dt * (div - 0.0);
    } else {
      v0 = 0.0;
    }
    if (row_int && int1) {
      const double div = (fx2 - fx1) / a.fc.dx + (fy1_p - fy1) / a.fc.dy;
      v1 = cur.s_now.y - a.dt * (div - 0.0);
    } else {
      v1 = 0.0;
    }
    const double g0 = cur.gam.x, g1 = cur.gam.y;
    if (own && ((!(row_int && int0) && g0 != 1.0) || (!(row_int && int1) && g1 != 1.0))) {
      // untouched by K1 and not overwritten by the relaxation: keep what the storage holds
      const double2 old = ld2o(a.s_new.p, o + col);
      if (!(row_int && int0) && g0 != 1.0) v0 = old.x;
      if (!(row_int && int1) && g1 != 1.0) v1 = old.y;
    }
    if (own && (g0 != 0.0 || g1 != 0.0)) {
      const double2 ref = ld2o(a.s_ref.p, o + col);
      if (g0 != 0.0) v0 = relax_point(g0, v0, ref.x);
      if (g1 != 0.0) v1 = relax_point(g1, v1, ref.y);
    }
    if (out1) {
      *reinterpret_cast<double2 *>(reinterpret_cast<char *>(a.spre.p) + (o + col)) = make_double2(v0, v1);
    } else if (out0) {
      sto(a.spre.p, o + col, v0);
    }
    fy0 = fy0_p;
    fy1 = fy1_p;
    o += row;
    og += grow;
  }
}

// ---------------------------------------------------------------- kernel T (moist stage)
// The water constituents of one RK stage in one pass, between kernels A and B.  The reference
// (dycore.py:L762-L812 around rk3ws_si.py:L126-L175) runs, per constituent: density
// (sq = s q, clipped; now and int: dwarfs/diagnostics.py:L400-L416), its share of K1
// (prognostics/utils.py:L101-L134), mass_fraction (q = sq_new / s_new, clipped; L434-L450) and the
// lateral relaxation (enforce_raw) -- 4 + 1/3 launches and 10 full-field passes per constituent.
// Here one thread per point forms the products s_int q_int on the stencil's cross on the fly
// (the three constituents share the loads of s_int and the face velocities), steps, divides by
// the stage's s (after its first relaxation: scratch_s, which is why this kernel runs before the
// momentum kernel updates s in place) and relaxes: reads s_now, s_int, s_pre, 2 x 3 q (+ su_int,
// sv_int or u, v), writes 3 q = 14 words per point, against 37 in the per-stencil path.
// Same operations in the same order, hence the same bits.  Requirement as for kernel MV:
// gamma == 1 on the nb outermost rings (the stale sq_new the reference divides there never
// survives the relaxation).
template <int SCHEME, bool DERIVE>
__global__ void __launch_bounds__(256) stage_tracers_kernel(const StageArgs a) {
  using F = Flux<SCHEME>;
  constexpr int E = F::extent;
  constexpr int NW = 2 * E;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= a.nx || j >= a.ny) return;
  const int nb = a.nb;
  const bool interior = i >= nb && i < a.nx - nb && j >= nb && j < a.ny - nb;
  const double gam = a.gamma.ld(i, j, 0);
  const long long sj = a.s_int.s1;
  const long long o = i + j * sj + (long long)k * a.s_int.s2;  // same geometry for every 3-D field
  if (!interior) {
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      if (t < a.ntr) {
        const double old = gam == 1.0 ? 0.0 : a.q_new[t].p[o];  // untouched by the step; the relaxation decides
        a.q_new[t].p[o] = relax_point(gam, old, __ldg(a.q_ref[t].p + o));
      }
    }
    return;
  }
  // s_int on the cross: sx[n] = s_int(i - E + n, j), sy[n] = s_int(i, j - E + n), n = 0 .. 2E
  double sx[NW + 1], sy[NW + 1];
  const double *ps = a.s_int.p + o;
#pragma unroll
  for (int n = 0; n <= NW; ++n) {
    sx[n] = __ldg(ps + (n - E));
    sy[n] = n == E ? sx[E] : __ldg(ps + (n - E) * sj);
  }
  double uq_m, uq_p, vq_m, vq_p;
  if (DERIVE) {  // velocity_x / velocity_y, dwarfs/diagnostics.py:L219-L272
    const double *pu = a.su_int.p + o, *pv = a.sv_int.p + o;
    const double su_c = __ldg(pu), sv_c = __ldg(pv);
    uq_m = F::prep(qdiv(__ldg(pu - 1) + su_c, sx[E - 1] + sx[E]), a.fc);
    uq_p = F::prep(qdiv(su_c + __ldg(pu + 1), sx[E] + sx[E + 1]), a.fc);
    vq_m = F::prep(qdiv(__ldg(pv - sj) + sv_c, sy[E - 1] + sy[E]), a.fc);
    vq_p = F::prep(qdiv(sv_c + __ldg(pv + sj), sy[E] + sy[E + 1]), a.fc);
  } else {
    uq_m = F::prep(__ldg(a.u_int.p + o), a.fc);
    uq_p = F::prep(__ldg(a.u_int.p + o + 1), a.fc);
    vq_m = F::prep(__ldg(a.v_int.p + o), a.fc);
    vq_p = F::prep(__ldg(a.v_int.p + o + sj), a.fc);
  }
  const double s_now = __ldg(a.s_now.p + o), s_pre = __ldg(a.spre.p + o);
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    if (t < a.ntr) {
      const double *pq = a.q_int[t].p + o;
      double qx[NW + 1], qy[NW + 1];
#pragma unroll
      for (int n = 0; n <= NW; ++n) {  // density of the constituent, clipped (diagnostics.py:L400-L416)
        const double x = sx[n] * __ldg(pq + (n - E));
        qx[n] = x > 0.0 ? x : 0.0;
        if (n == E) {
          qy[n] = qx[n];
        } else {
          const double y = sy[n] * __ldg(pq + (n - E) * sj);
          qy[n] = y > 0.0 ? y : 0.0;
        }
      }
      const double fxm = F::eval_v(uq_m, qx), fxp = F::eval_v(uq_p, qx + 1);
      const double fym = F::eval_v(vq_m, qy), fyp = F::eval_v(vq_p, qy + 1);
      const double div = (fxp - fxm) / a.fc.dx + (fyp - fym) / a.fc.dy;
      double sq_now = s_now * __ldg(a.q_now[t].p + o);
      sq_now = sq_now > 0.0 ? sq_now : 0.0;
      const double sq_new = sq_now - a.dt * (div - 0.0);  // prognostics/utils.py:L101-L134
      double q = qdiv(sq_new, s_pre);                      // diagnostics.py:L434-L450
      q = q > 0.0 ? q : 0.0;
      if (gam != 0.0) q = relax_point(gam, q, __ldg(a.q_ref[t].p + o));
      a.q_new[t].p[o] = q;
    }
  }
}

// Kernel T, tiled (default; TB200_T_IMPL=point selects the kernel above).  The thread-per-point
// version evaluates every face flux twice, every product s q up to 2E+1 times and derives four
// face velocities per point: 380 us per launch at configs[2] (21 % of the step, launch list
// r02e_launches_c3.csv) where the traffic would allow 90.  Here a block owns a tile of 31 x 16
// points of one level:
//   phase 1  the densities clip(s_int q_int) of the three constituents on the tile + E halo
//            points a side go to shared memory, each formed once (coalesced row loads);
//   phase 2  thread (tx, ty) evaluates the flux through the LEFT face and through the BOTTOM face
//            of its point for all three constituents, with one velocity diagnosis per face; lane
//            31 of a row and row 16 of the block are lenders (their faces close the tile, they own
//            no point), so every face of the tile is evaluated exactly once; x-fluxes reach the
//            left neighbour by shuffle, y-fluxes through shared memory;
//   phase 3  divergence, step, division by the stage's s, clipping, relaxation, store.
// Same point formulas in the same order as the kernel above, hence the same bits.
// Tile height TT_Y (TB200_T_ROWS=4|8|16, default 8): the phases of a block are separated by
// barriers, so loads and arithmetic only overlap ACROSS blocks; 16-row tiles (544 threads, two
// blocks per SM) ran at 29 % of the DRAM bandwidth (profiles/README.md, round 2), shorter tiles
// trade y-halo work for more resident blocks.
constexpr int TT_X = 31;
template <int SCHEME, bool DERIVE, int TT_Y>
__global__ void __launch_bounds__(32 * (TT_Y + 1)) stage_tracers_tile_kernel(const StageArgs a) {
  using F = Flux<SCHEME>;
  constexpr int E = F::extent;
  constexpr int NW = 2 * E;
  constexpr int SXW = 32 + NW, SYW = TT_Y + NW;  // cells i0 - E .. i0 + 31 + E - 1 (one spare), j0 - E .. j0 + 15 + E
  __shared__ double sq[3][SYW][SXW + 1];
  __shared__ double fys[3][TT_Y + 1][32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i0 = blockIdx.x * TT_X, j0 = blockIdx.y * TT_Y, k = blockIdx.z;
  const int nx = a.nx, ny = a.ny, nb = a.nb;
  const long long sj = a.s_int.s1;
  const long long pl = (long long)k * a.s_int.s2;
  // ---- phase 1: densities on the tile + halo (indices clamped into the storage: clamped
  // values only reach faces of points outside the computational domain)
  for (int n = ty * 32 + tx; n < SXW * SYW; n += 32 * (TT_Y + 1)) {
    const int ly = n / SXW, lx = n - ly * SXW;
    const int gi = min(max(i0 - E + lx, 0), nx - 1), gj = min(max(j0 - E + ly, 0), ny - 1);
    const long long o = pl + gi + gj * sj;
    const double sv_ = __ldg(a.s_int.p + o);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const double x = sv_ * __ldg(a.q_int[t].p + o);  // diagnostics.py:L400-L416
      sq[t][ly][lx] = x > 0.0 ? x : 0.0;
    }
  }
  __syncthreads();
  // ---- phase 2: fluxes through the left and the bottom face of point (i, j)
  const int i = i0 + tx, j = j0 + ty;
  const int ic = min(i, nx - 1), jc = min(j, ny - 1);  // lenders / edge threads stay inside the storage
  const long long o = pl + ic + jc * sj;
  double uq, vq;
  if (DERIVE) {  // velocity_x / velocity_y, dwarfs/diagnostics.py:L219-L272
    const long long ol = pl + max(ic - 1, 0) + jc * sj, ob = pl + ic + max(jc - 1, 0) * sj;
    const double s_c = __ldg(a.s_int.p + o);
    uq = F::prep(qdiv(__ldg(a.su_int.p + ol) + __ldg(a.su_int.p + o), __ldg(a.s_int.p + ol) + s_c), a.fc);
    vq = F::prep(qdiv(__ldg(a.sv_int.p + ob) + __ldg(a.sv_int.p + o), __ldg(a.s_int.p + ob) + s_c), a.fc);
  } else {
    uq = F::prep(__ldg(a.u_int.p + o), a.fc);
    vq = F::prep(__ldg(a.v_int.p + o), a.fc);
  }
  double fx[3], fxp[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    // face i: cells i - E .. i + E - 1 = shared columns tx .. tx + NW - 1 of row ty + E
    double wx[NW], wy[NW];
    const int yr = min(ty + E, SYW - 1);  // (the lender row's x-fluxes are not used)
#pragma unroll
    for (int m = 0; m < NW; ++m) {
      wx[m] = sq[t][yr][tx + m];
      wy[m] = sq[t][ty + m][tx + E];      // face j: cells j - E .. j + E - 1 = shared rows ty .. ty + NW - 1
    }
    fx[t] = F::eval_v(uq, wx);
    fys[t][ty][tx] = F::eval_v(vq, wy);
    fxp[t] = __shfl_down_sync(0xffffffffu, fx[t], 1);
  }
  __syncthreads();
  // ---- phase 3: the point update
  if (tx >= TT_X || ty >= TT_Y || i >= nx || j >= ny) return;
  const bool interior = i >= nb && i < nx - nb && j >= nb && j < ny - nb;
  const double gam = a.gamma.ld(i, j, 0);
  if (!interior) {
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const double old = gam == 1.0 ? 0.0 : a.q_new[t].p[o];  // untouched by the step; the relaxation decides
      a.q_new[t].p[o] = relax_point(gam, old, __ldg(a.q_ref[t].p + o));
    }
    return;
  }
  const double s_now = __ldg(a.s_now.p + o), s_pre = __ldg(a.spre.p + o);
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const double div = (fxp[t] - fx[t]) / a.fc.dx + (fys[t][ty + 1][tx] - fys[t][ty][tx]) / a.fc.dy;
    double sq_now = s_now * __ldg(a.q_now[t].p + o);
    sq_now = sq_now > 0.0 ? sq_now : 0.0;
    const double sq_new = sq_now - a.dt * (div - 0.0);  // prognostics/utils.py:L101-L134
    double q = qdiv(sq_new, s_pre);                      // diagnostics.py:L434-L450
    q = q > 0.0 ? q : 0.0;
    if (gam != 0.0) q = relax_point(gam, q, __ldg(a.q_ref[t].p + o));
    a.q_new[t].p[o] = q;
  }
}

int t_impl() {  // TB200_T_IMPL=point|tile
  static int impl = -1;
  if (impl < 0) {
    const char *e = getenv("TB200_T_IMPL");
    impl = (e != nullptr && strcmp(e, "point") == 0) ? 0 : 1;
  }
  return impl;
}

template <int SCHEME>
int launch_tracers(const StageArgs &a, cudaStream_t st) {
  if (t_impl() == 1) {
    static int rows = -1;
    if (rows < 0) {
      const char *e = getenv("TB200_T_ROWS");
      rows = e == nullptr ? 8 : atoi(e);
      if (rows != 4 && rows != 8 && rows != 16) rows = 8;
    }
    dim3 block(32, rows + 1, 1);
    dim3 grid((a.nx + TT_X - 1) / TT_X, (a.ny + rows - 1) / rows, a.nz);
#define TB200_T_LAUNCH(R)                                                     \
    if (a.derive_uv)                                                          \
      stage_tracers_tile_kernel<SCHEME, true, R><<<grid, block, 0, st>>>(a);  \
    else                                                                      \
      stage_tracers_tile_kernel<SCHEME, false, R><<<grid, block, 0, st>>>(a);
    if (rows == 4) { TB200_T_LAUNCH(4) } else if (rows == 16) { TB200_T_LAUNCH(16) } else { TB200_T_LAUNCH(8) }
#undef TB200_T_LAUNCH
    return check_launch("isentropic_stage_moist/T(tile)");
  }
  dim3 block(64, 4, 1);
  dim3 grid((a.nx + 63) / 64, (a.ny + 3) / 4, a.nz);
  if (a.derive_uv)
    stage_tracers_kernel<SCHEME, true><<<grid, block, 0, st>>>(a);
  else
    stage_tracers_kernel<SCHEME, false><<<grid, block, 0, st>>>(a);
  return check_launch("isentropic_stage_moist/T");
}

// The Exner function of eight levels as ONE real call: kernel B is unrolled over the levels, so
// an inlined copy of the power per level would not fit the instruction cache, and the eight
// evaluations are interleaved instruction by instruction (pow_pos_n): with two resident warps
// per scheduler (kernel B keeps a column in registers) eight interleaved chains cover the
// 8-cycle DFMA latency from within one warp (experiments/fp64_latency.cu).
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

// Kernel B for SMALL grids.  One thread per column is the right shape when there are a million
// columns (config 5); a 161 x 161 grid has 26 000, i.e. 1.4 warps per scheduler, each walking
// through ~4000 dependent fp64 instructions: 45 us where the data would allow 5 (profiles/
// README.md, round 1d).  Here a block of 32 columns x 8 threads splits the column work by what is
// actually serial: only the two running sums are (one add per level each); the divisions by pref,
// the Exner powers (eight levels per thread -- exactly one exner8 call) and the products entering
// the sums are evaluated by all 256 threads out of shared memory.  Every value is produced by the
// same operation sequence as in stage_b_kernel: bit-identical results.
template <int NZC>
__global__ void __launch_bounds__(256) stage_b_coop_kernel(const StageArgs a) {
  static_assert(NZC % 8 == 0 && NZC <= 64, "eight threads per column, eight levels each");
  __shared__ double sm[NZC][33];
  const int tx = threadIdx.x, ty = threadIdx.y;  // column within the block / level group
  const int i = blockIdx.x * 32 + tx, j = blockIdx.y;
  const bool live = i < a.nx;
  const int ic = live ? i : a.nx - 1;
  const int nz = a.nz;
  const double kappa = a.rd / a.cp;
  const double gdz = a.g * a.dz;
  // increments of the pressure sum, levels ty, ty + 8, ... (rows of 32 columns: coalesced)
  {
    const double *ps = a.spre.p + (ic + j * a.spre.s1);
    const long long s2 = a.spre.s2;
#pragma unroll
    for (int k = ty; k < NZC; k += 8) sm[k][tx] = k < nz ? gdz * __ldg(ps + k * s2) : 0.0;
  }
  __syncthreads();
  if (ty == 0) {  // pressure at interface k + 1, diagnostics.py:L425-L428
    double p = a.pt;
#pragma unroll
    for (int k = 0; k < NZC; ++k) {
      if (k < nz) {
        p = p + sm[k][tx];
        sm[k][tx] = p;
      }
    }
  }
  __syncthreads();
  {  // Exner function of the eight interfaces 8 ty + 1 .. 8 ty + 8 and their weight in the sum
    D8 x;
#pragma unroll
    for (int n = 0; n < 8; ++n) x.v[n] = 8 * ty + n < nz ? sm[8 * ty + n][tx] / a.cpref : 1.0;
    if (8 * ty < nz) {
      const D8 r = exner8(x, kappa, a.cp);
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int k = 8 * ty + n;
        // level nz - 1 keeps the Exner value itself (surface term), the others dz * exn
        sm[k][tx] = k == nz - 1 ? r.v[n] : a.dz * r.v[n];
      }
    }
  }
  __syncthreads();
  if (ty == 0 && live) {  // upward sweep, diagnostics.py:L433-L438
    double *pm = a.mtg.p + (i + j * a.mtg.s1);
    const long long m2 = a.mtg.s2;
    const double ex_s = sm[nz - 1][tx];
    const double mtg_s = a.theta_s * ex_s + a.g * a.hs.ld(i, j, 0);
    double m = mtg_s + 0.5 * a.dz * ex_s;
    pm[(nz - 1) * m2] = m;
    for (int k = nz - 2; k >= 0; --k) {
      m = m + sm[k][tx];
      sm[k][tx] = m;
    }
  }
  __syncthreads();
  if (live) {  // store the levels below the surface one, rows of 32 columns
    double *pm = a.mtg.p + (i + j * a.mtg.s1);
    const long long m2 = a.mtg.s2;
#pragma unroll
    for (int k = ty; k < NZC; k += 8)
      if (k < nz - 1) pm[k * m2] = sm[k][tx];
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

// ---------------------------------------------------------------- kernel MV, two columns per lane
// The ring kernel above still issues ~445 instructions per warp-row of 30 points, of which only
// 141 are fp64: every global access costs two integer instructions (64-bit address formation;
// sm_100a has no [uniform base + 32-bit offset] global addressing), every neighbour one LDS and
// every shared face one shuffle.  Here
//   * a lane owns TWO adjacent columns (c0 even, c1 = c0 + 1): all global accesses are 16-byte
//     LDG.128 / STG.128 (half the address arithmetic per point), the y-window and the
//     x-neighbours are LDS.128, the face between c0 and c1 needs no shuffle, and a warp carries
//     two independent dependency chains through every formula;
//   * the rows with reuse (su_int, sv_int, mtg_now, mtg_new) go global -> shared memory with
//     cp.async (LDGSTS, 16 bytes per lane, no staging registers, no STS), one row ahead;
//   * the reference fields of the relaxation band / damping layer are requested one row ahead
//     like every other own-column value (their DRAM latency used to sit on every row there).
// One warp = 64 columns of one level, 60 of them owned (lanes 1..30); lane 0 recomputes the two
// columns to the left (their final s, su feed lane 1's u), lane 31 lends its left face.  A
// block = WX x WY warps (WX side by side, WY strips of LJ rows).
// Ring row: grid columns c_first-6 .. c_first+65 (72 doubles, 16-byte aligned pairs); su / sv
// rings of 8 rows, mtg rings of 4 rows, slot = row & 7 / & 3 relative to the strip.
// Arithmetic and results are bit-identical to the other momentum kernels.
// Requirements (checked on the host, else the ring kernel runs): 16-byte aligned bases, even
// row / plane pitches, row pitch >= nx + 1.
constexpr int MV2_COLS = 60;
constexpr int RW2 = 72;
constexpr int SU_RING2 = 8, MT_RING2 = 4;
constexpr int REF_RING2 = 2;  // rows r (in use) and r+1 (in flight) of s_ref, su_ref, sv_ref
constexpr int WARP_DOUBLES2 = (2 * SU_RING2 + 2 * MT_RING2) * RW2 + 3 * REF_RING2 * 64;
constexpr int MV2_BLOCK_DEFAULT = 1, MV2_DEF_WX = 3, MV2_DEF_WY = 1;  // TB200_MV_BLOCK: 0 = 2x2, 1 = 3x1, 2 = 6x1
constexpr int MV2_PF = 2;  // rows of DRAM -> L2 prefetch ahead of the loads (measured: 0 -> 2.00 ms, 2 -> 1.79, 3 -> 1.82, 6 -> 2.22)

__device__ __forceinline__ double2 ldo2(const double *base, unsigned off) {
  return __ldg(reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(base) + off));
}
__device__ __forceinline__ void sto2(double *base, unsigned off, double2 v) {
  *reinterpret_cast<double2 *>(reinterpret_cast<char *>(base) + off) = v;
}
__device__ __forceinline__ double2 lds2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
// 16 bytes global -> shared, asynchronously (LDGSTS, L2 only)
__device__ __forceinline__ void cp_async16(unsigned smem_dst, const double *base, unsigned off) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst),
               "l"(reinterpret_cast<const char *>(base) + off)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct OwnLoads2 {
  double2 v_n, u_c;  // v_int at row r+1, u_int at row r (not DERIVE)
  double2 s_i2;      // s_int at row r+2 (DERIVE)
  double2 s_pre, s_now, su_now, sv_now;
};

// Template flags (tb200_isentropic_stage.derive_uv_in / .skip_uv_out):
//   DERIVE  the advecting velocities are re-diagnosed from s_int and the su_int / sv_int rows
//           already in the rings (velocity_x / velocity_y, dwarfs/diagnostics.py:L219-L272)
//           instead of read from u_int / v_int: one stream (s_int) instead of two;
//   UVOUT   u_new, v_new are diagnosed and written.  Without it (an intermediate stage) lane 0
//           has nothing to recompute, the warm-up row disappears, and -- what matters most --
//           only the 30 owner lanes touch the streams without reuse, i.e. exactly the 15
//           32-byte sectors of the warp's 60 columns: the two-column shift of lane 0 used to cost
//           two more sectors per row and stream (17 / 15), because the x-neighbour warp runs far
//           enough away in time for the shared sectors to have left L2 (profiles/README.md,
//           round 2).  For the same reason the halo pairs nobody reads are not fetched any more.
// WX x WY warps per block: WX side by side, WY strips of LJ rows.
//   TND     slow tendencies of su, sv are passed (prognostics/utils.py:L191-L204); the reference's
//           `- 0.0` becomes `- tnd`.  A separate instantiation, so that the kernels of a run
//           without tendencies (the benchmark configurations) are untouched.
template <int SCHEME, int LJ, int WX, int WY, bool DERIVE, bool UVOUT, bool TND = false>
__global__ void __launch_bounds__(32 * WX * WY, 384 / (32 * WX * WY)) stage_mv2_kernel(const StageArgs a) {
  using F = Flux<SCHEME>;
  constexpr int E = F::extent;
  constexpr int NW = 2 * E;
  extern __shared__ __align__(16) double ring_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int xw = (blockIdx.x + a.bx0) * WX + warp % WX;
  const int j0 = ((blockIdx.y + a.by0) * WY + warp / WX) * LJ;
  if (xw * MV2_COLS >= a.nx || j0 >= a.ny) return;  // warp-uniform (no block-wide barrier in this kernel)
  double *r_su = ring_smem + warp * WARP_DOUBLES2;
  const int c0 = xw * MV2_COLS + 2 * lane - 2;  // first of the lane's two columns (even)
  const int jend = min(j0 + LJ, a.ny);
  const int k = blockIdx.z;
  const int nx = a.nx, ny = a.ny, nb = a.nb;

  const bool own_lane = lane >= 1 && lane <= MV2_COLS / 2;
  const bool out0 = own_lane && c0 < nx, out1 = own_lane && c0 + 1 < nx;
  const bool col_int0 = c0 >= nb && c0 < nx - nb, col_int1 = c0 + 1 >= nb && c0 + 1 < nx - nb;
  const int pmax = (nx - 1) & ~1;        // last pair start inside the row
  const int cm = min(max(c0, 0), pmax);  // own pair, clamped into the row
  // who reads the streams without reuse: the owners, plus lane 0 when it recomputes the two
  // columns to the left for u_new; u_int also feeds lane 31's left face
  const bool ld_own = own_lane || (UVOUT && lane == 0);
  const bool ld_u = lane > 0 || UVOUT;
  // halo pairs of the su / sv rings: ring columns 2-3 (lane 1) and 68-69 (lane 2); of the mtg
  // rings: ring columns 2-3, only for lane 0's recomputation.  Ring columns 0-1 and 70-71 feed
  // nothing that is used (lane 0's left face); they are zeroed once.
  const bool halo_su = lane == 1 || lane == 2;
  const bool halo_mt = UVOUT && lane == 1;
  const int th = lane == 1 ? 2 : 68;
  const int ch = min(max(xw * MV2_COLS - 6 + th, 0), pmax);
  const int t = 2 * lane + 4;  // ring column of c0

  const unsigned row = (unsigned)a.s_now.s1 * 8u;
  const unsigned plane = (unsigned)k * (unsigned)a.s_now.s2 * 8u;
  const unsigned grow = (unsigned)a.gamma.s1 * 8u;
  const double r_damp = a.damp ? a.rmat.ld(0, 0, k) : 0.0;
  const double one_m_eps = 1.0 - a.eps;
  // first row computed: the warm-up row j0-1 seeds the v_new diagnosis
  const int r0 = (UVOUT && j0 > 0) ? j0 - 1 : j0;
  const double2 zero2 = make_double2(0.0, 0.0);
  // s is updated in place (scratch_s is s_new): stored only where relaxation / damping change it
  const bool s_inplace = !UVOUT && a.spre.p == a.s_new.p;

  // shared-memory byte addresses of this lane's own / halo entry of ring row 0
  constexpr unsigned SLOT = RW2 * 8u;                       // bytes per ring row
  constexpr unsigned SV_OFF = SU_RING2 * SLOT;              // sv ring relative to su ring
  constexpr unsigned MN_OFF = 2 * SU_RING2 * SLOT;          // mtg_now ring
  constexpr unsigned MW_OFF = MN_OFF + MT_RING2 * SLOT;     // mtg_new ring
  const unsigned sm_own = (unsigned)__cvta_generic_to_shared(r_su + t);
  const unsigned sm_halo = (unsigned)__cvta_generic_to_shared(r_su + th);
  // reference fields: [stream][slot][64 columns] behind the rings, own pair at 2 * lane
  const double *w_ref = r_su + (2 * SU_RING2 + 2 * MT_RING2) * RW2 + 2 * lane;
  const unsigned sm_ref = (unsigned)__cvta_generic_to_shared(w_ref);
  const unsigned col_o = (unsigned)cm * 8u, col_h = (unsigned)ch * 8u;

  if (lane < 2 * SU_RING2 + 2 * MT_RING2) {  // ring columns nobody fetches (ring row = lane)
    *reinterpret_cast<double2 *>(r_su + lane * RW2) = zero2;
    *reinterpret_cast<double2 *>(r_su + lane * RW2 + 70) = zero2;
    if (lane >= 2 * SU_RING2) {
      *reinterpret_cast<double2 *>(r_su + lane * RW2 + 68) = zero2;
      if (!UVOUT) *reinterpret_cast<double2 *>(r_su + lane * RW2 + 2) = zero2;
    }
  }
  __syncwarp();

  // grid row rr of su_int / sv_int -> ring slot q_su & 7; of mtg_now / mtg_new -> slot q_mt & 3
  auto fetch_su = [&](int q_su, int rr) {
    const unsigned o = plane + (unsigned)max(rr, 0) * row, d = (unsigned)(q_su & (SU_RING2 - 1)) * SLOT;
    cp_async16(sm_own + d, a.su_int.p, o + col_o);
    cp_async16(sm_own + d + SV_OFF, a.sv_int.p, o + col_o);
    if (halo_su) {
      cp_async16(sm_halo + d, a.su_int.p, o + col_h);
      cp_async16(sm_halo + d + SV_OFF, a.sv_int.p, o + col_h);
    }
  };
  auto fetch_mt = [&](int q_mt, int rr) {
    const unsigned o = plane + (unsigned)max(rr, 0) * row, d = (unsigned)(q_mt & (MT_RING2 - 1)) * SLOT;
    cp_async16(sm_own + d + MN_OFF, a.mtg_now.p, o + col_o);
    cp_async16(sm_own + d + MW_OFF, a.mtg.p, o + col_o);
    if (halo_mt) {
      cp_async16(sm_halo + d + MN_OFF, a.mtg_now.p, o + col_h);
      cp_async16(sm_halo + d + MW_OFF, a.mtg.p, o + col_h);
    }
  };

  // ---- prologue: su / sv rows r0-E .. r0+E-1 (slots 0 .. 2E-1), mtg rows r0-1, r0 (slots 0, 1),
  // then the rows the first iteration waits for (r0+E -> slot 2E, r0+1 -> slot 2) as the
  // youngest group.  Rows up to jend + E are touched: inside the allocation (>= nz + 1 planes,
  // checked on the host).
#pragma unroll
  for (int m = 0; m < NW; ++m) fetch_su(m, r0 - E + m);
  fetch_mt(0, r0 - 1);
  fetch_mt(1, r0);
  cp_async_commit();
  unsigned o_c = plane + (unsigned)r0 * row + col_o;  // own pair, row r

  // The streams without reuse are loaded UNCONDITIONALLY, one row ahead of their use: a lane that
  // does not need a stream (ld_own / ld_u false) reads the pair of the owner lane next to it --
  // a sector that lane fetches anyway, so no extra traffic -- and discards the value.  A
  // predicated load with a zero alternative (`p ? load : 0`) compiles to a branch whose merge
  // moves CONSUME the load right after its issue: the row-ahead request then stalls the warp for
  // a full memory latency on every row (profiles/README.md, round 2: 13 % of all stall samples on
  // one IMAD.MOV).
  const unsigned d_own = ld_own ? 0u : (unsigned)(min(max(c0 + (lane == 0 ? 2 : -2), 0), pmax) - cm) * 8u;
  const unsigned d_u = ld_u ? 0u : d_own;
  auto load_own = [&](unsigned oc) {
    OwnLoads2 L;
    if (DERIVE) {
      L.v_n = L.u_c = zero2;
      L.s_i2 = ldo2(a.s_int.p, oc + 2 * row);  // every lane: lane l + 1 takes s_int[c0 - 1] from lane l
    } else {
      L.v_n = ldo2(a.v_int.p, oc + d_own + row);
      L.u_c = ldo2(a.u_int.p, oc + d_u);
      L.s_i2 = zero2;
    }
    L.s_pre = ldo2(a.spre.p, oc + d_own);
    L.s_now = ldo2(a.s_now.p, oc + d_own);
    L.su_now = ldo2(a.su_now.p, oc + d_own);
    L.sv_now = ldo2(a.sv_now.p, oc + d_own);
    return L;
  };
  // The reference fields of the relaxation band / damping layer are requested one row ahead of
  // their use like everything else (cp.async into a two-row ring; the relaxation coefficients
  // run two rows ahead to decide where).  Row rr -> slot rr & 1.
  auto fetch_ref = [&](int rr, unsigned oc, double2 gam) {
    if (ld_own && (gam.x != 0.0 || gam.y != 0.0 || r_damp != 0.0)) {
      const unsigned d = sm_ref + (unsigned)(rr & 1) * 512u;
      cp_async16(d, a.s_ref.p, oc);
      cp_async16(d + REF_RING2 * 512u, a.su_ref.p, oc);
      cp_async16(d + 2 * REF_RING2 * 512u, a.sv_ref.p, oc);
    }
  };
  auto load_gamma = [&](int rr) {  // rows beyond the domain repeat the last one (never used)
    return ldo2(a.gamma.p, (unsigned)min(rr, ny - 1) * grow + col_o);
  };
  double2 gam_a = load_gamma(r0), gam_b = load_gamma(r0 + 1);  // rows r and r+1
  // DERIVE: s_int rows r (si_a) and r+1 (si_b) ride along in registers; the prologue also needs
  // row r0-1 for the y-face r0
  double2 si_a = zero2, si_b = zero2, v_first = zero2;
  if (DERIVE) {
    v_first = ldo2(a.s_int.p, plane + (unsigned)max(r0 - 1, 0) * row + col_o);  // s_int at row r0-1
    si_a = ldo2(a.s_int.p, o_c);
    si_b = ldo2(a.s_int.p, o_c + row);
  } else {
    v_first = ldo2(a.v_int.p, o_c + d_own);
  }
  OwnLoads2 nxt = load_own(o_c);
  fetch_su(NW, r0 + E);
  fetch_mt(2, r0 + 1);
  fetch_ref(r0, o_c, gam_a);
  cp_async_commit();

  cp_async_wait<1>();
  __syncwarp();
  double fy_su0, fy_su1, fy_sv0, fy_sv1;
  {
    double ysu0[NW], ysu1[NW], ysv0[NW], ysv1[NW];
#pragma unroll
    for (int m = 0; m < NW; ++m) {
      const double2 vu = lds2(r_su + m * RW2 + t), vv = lds2(r_su + (SU_RING2 + m) * RW2 + t);
      ysu0[m] = vu.x; ysu1[m] = vu.y; ysv0[m] = vv.x; ysv1[m] = vv.y;
    }
    double vq0, vq1;
    if (DERIVE) {  // window entry m holds row r0 - E + m: rows r0-1, r0 are entries E-1, E
      vq0 = F::prep(qdiv(ysv0[E - 1] + ysv0[E], v_first.x + si_a.x), a.fc);
      vq1 = F::prep(qdiv(ysv1[E - 1] + ysv1[E], v_first.y + si_a.y), a.fc);
    } else {
      vq0 = F::prep(v_first.x, a.fc);
      vq1 = F::prep(v_first.y, a.fc);
    }
    fy_su0 = F::eval_v(vq0, ysu0); fy_su1 = F::eval_v(vq1, ysu1);
    fy_sv0 = F::eval_v(vq0, ysv0); fy_sv1 = F::eval_v(vq1, ysv1);
  }
  double sv_prev0 = 0.0, sv_prev1 = 0.0, s_prev0 = 0.0, s_prev1 = 0.0;

  // DRAM -> L2 MV2_PF rows ahead with two instructions per row for all streams: lane
  // 10 l + f takes the 128-byte lines 2 l and 2 l + 1 of stream f, counted from the line that
  // holds the warp's first owned column, as far as they overlap the owned columns (the halo
  // sectors beyond are left to the loads: prefetching whole neighbour lines is what the
  // neighbour warp, far away in time, would do again)
  const double *pf_base;
  unsigned pf_off;
  bool pf_a, pf_b;
  {
    const int f = lane % 10, line = lane / 10;
    const double *bases[10] = {a.su_int.p, a.sv_int.p, DERIVE ? a.s_int.p : a.v_int.p,
                               DERIVE ? nullptr : a.u_int.p, a.spre.p,
                               a.s_now.p,  a.su_now.p, a.sv_now.p, a.mtg_now.p, a.mtg.p};
    const int ahead[10] = {E, E, DERIVE ? 2 : 1, 0, 0, 0, 0, 0, 1, 1};
    pf_base = bases[0];
    int ah = 0;
#pragma unroll
    for (int n = 0; n < 10; ++n)
      if (f == n) {
        pf_base = bases[n];
        ah = ahead[n];
      }
    const int b_lo = xw * MV2_COLS * 8, b_hi = min(b_lo + MV2_COLS * 8, (int)row);
    const int l0 = (b_lo & ~127) + 256 * line;
    pf_a = lane < 30 && pf_base != nullptr && l0 < b_hi;
    pf_b = lane < 30 && pf_base != nullptr && l0 + 128 < b_hi;
    if (pf_base == nullptr) pf_base = bases[0];
    pf_off = plane + (unsigned)(r0 + MV2_PF + ah) * row + (unsigned)l0;
  }

  const double *w_su = r_su + t;  // this lane's entry of su ring row 0
  for (int r = r0; r < jend; ++r) {
    const int q = r - r0;
    // rows r+E / r+1, requested one iteration ago, have landed; all lanes are done with row r-1
    cp_async_wait<0>();
    __syncwarp();
    fetch_su(q + NW + 1, r + E + 1);
    fetch_mt(q + 3, r + 2);
    fetch_ref(r + 1, o_c + row, gam_b);
    cp_async_commit();
    const OwnLoads2 cur = nxt;
    nxt = load_own(o_c + row);
    const double2 gam_c = load_gamma(r + 2);
    if (r + MV2_PF < jend) {
      if (pf_a) prefetch_l2(pf_base, pf_off);
      if (pf_b) prefetch_l2(pf_base, pf_off + 128u);
    }
    pf_off += row;

    // ---- y-faces r+1 of both columns: rows r-E+1 .. r+E sit in slots (q + 1 + m) & 7
    double ysu0[NW], ysu1[NW], ysv0[NW], ysv1[NW];
    const double *row_r = w_su;  // ring row of grid row r
#pragma unroll
    for (int m = 0; m < NW; ++m) {
      const double *pr = w_su + ((q + 1 + m) & (SU_RING2 - 1)) * RW2;
      if (m == E - 1) row_r = pr;
      const double2 vu = lds2(pr), vv = lds2(pr + SU_RING2 * RW2);
      ysu0[m] = vu.x; ysu1[m] = vu.y; ysv0[m] = vv.x; ysv1[m] = vv.y;
    }
    double vq0, vq1;
    if (DERIVE) {  // v at the y-face r+1: window entries E-1 (row r) and E (row r+1)
      vq0 = F::prep(qdiv(ysv0[E - 1] + ysv0[E], si_a.x + si_b.x), a.fc);
      vq1 = F::prep(qdiv(ysv1[E - 1] + ysv1[E], si_a.y + si_b.y), a.fc);
    } else {
      vq0 = F::prep(cur.v_n.x, a.fc);
      vq1 = F::prep(cur.v_n.y, a.fc);
    }
    const double fy_su0_p = F::eval_v(vq0, ysu0), fy_su1_p = F::eval_v(vq1, ysu1);
    const double fy_sv0_p = F::eval_v(vq0, ysv0), fy_sv1_p = F::eval_v(vq1, ysv1);

    // ---- x-faces at row r: left face of c0, face between c0 and c1; the right face of c1 is
    // the left face of the next lane's c0.  lsu[n] = su_int[c0 - E + n]
    double lsu[NW + 1], lsv[NW + 1];
    {
      const double *psu = row_r, *psv = row_r + SU_RING2 * RW2;
#pragma unroll
      for (int po = -4; po <= 2; po += 2) {  // aligned pairs (po, po + 1) around the own pair (0, 1)
        if (po == 0) continue;
        const bool need0 = po >= -E && po <= E, need1 = po + 1 >= -E && po + 1 <= E;
        if (need0 && need1) {
          const double2 vu = lds2(psu + po), vv = lds2(psv + po);
          lsu[po + E] = vu.x; lsu[po + 1 + E] = vu.y;
          lsv[po + E] = vv.x; lsv[po + 1 + E] = vv.y;
        } else if (need0) {
          lsu[po + E] = psu[po]; lsv[po + E] = psv[po];
        } else if (need1) {
          lsu[po + 1 + E] = psu[po + 1]; lsv[po + 1 + E] = psv[po + 1];
        }
      }
      lsu[E] = ysu0[E - 1]; lsu[E + 1] = ysu1[E - 1];
      lsv[E] = ysv0[E - 1]; lsv[E + 1] = ysv1[E - 1];
    }
    double uq0, uq1;
    if (DERIVE) {  // u at the left face of c0 and at the face between c0 and c1
      const double si_l = __shfl_up_sync(0xffffffffu, si_a.y, 1);  // s_int[c0 - 1] (lane 0: unused face)
      uq0 = F::prep(qdiv(lsu[E - 1] + lsu[E], si_l + si_a.x), a.fc);
      uq1 = F::prep(qdiv(lsu[E] + lsu[E + 1], si_a.x + si_a.y), a.fc);
    } else {
      uq0 = F::prep(cur.u_c.x, a.fc);
      uq1 = F::prep(cur.u_c.y, a.fc);
    }
    const double fx_su0 = F::eval_v(uq0, lsu), fx_su1 = F::eval_v(uq1, lsu + 1);
    const double fx_sv0 = F::eval_v(uq0, lsv), fx_sv1 = F::eval_v(uq1, lsv + 1);
    const double fx_su2 = __shfl_down_sync(0xffffffffu, fx_su0, 1);
    const double fx_sv2 = __shfl_down_sync(0xffffffffu, fx_sv0, 1);

    // ---- point updates (prognostics/utils.py:L191-L204)
    const bool row_int = r >= nb && r < ny - nb;
    const bool int0 = col_int0 && row_int, int1 = col_int1 && row_int;
    double s0 = cur.s_pre.x, s1 = cur.s_pre.y, su0 = 0.0, su1 = 0.0, sv0 = 0.0, sv1 = 0.0;
    if (int0 || int1) {
      // slow tendencies of the own pair (loaded at their use: not the benchmark's path)
      const double2 tnd_su = TND ? ldo2(a.su_tnd.p, o_c) : zero2, tnd_sv = TND ? ldo2(a.sv_tnd.p, o_c) : zero2;
      // Montgomery potential: rows r-1, r, r+1 in slots (q + m) & 3
      const double *bm_m = w_su + (2 * SU_RING2 + (q & (MT_RING2 - 1))) * RW2;
      const double *bm_c = w_su + (2 * SU_RING2 + ((q + 1) & (MT_RING2 - 1))) * RW2;
      const double *bm_p = w_su + (2 * SU_RING2 + ((q + 2) & (MT_RING2 - 1))) * RW2;
      constexpr int MW = MT_RING2 * RW2;  // mtg_new relative to mtg_now
      const double2 mn_c = lds2(bm_c), mw_c = lds2(bm_c + MW);
      const double2 mn_m = lds2(bm_m), mw_m = lds2(bm_m + MW);
      const double2 mn_p = lds2(bm_p), mw_p = lds2(bm_p + MW);
      const double mn_l = bm_c[-1], mw_l = bm_c[MW - 1], mn_r = bm_c[2], mw_r = bm_c[MW + 2];
      {
        const double div = (fx_su1 - fx_su0) / a.fc.dx + (fy_su0_p - fy_su0) / a.fc.dy;
        const double pg_now = one_m_eps * cur.s_now.x * (mn_c.y - mn_l) / a.two_dx;
        const double pg_new = a.eps * s0 * (mw_c.y - mw_l) / a.two_dx;
        su0 = cur.su_now.x - a.dt * (div + pg_now + pg_new - tnd_su.x);
      }
      {
        const double div = (fx_su2 - fx_su1) / a.fc.dx + (fy_su1_p - fy_su1) / a.fc.dy;
        const double pg_now = one_m_eps * cur.s_now.y * (mn_r - mn_c.x) / a.two_dx;
        const double pg_new = a.eps * s1 * (mw_r - mw_c.x) / a.two_dx;
        su1 = cur.su_now.y - a.dt * (div + pg_now + pg_new - tnd_su.y);
      }
      {
        const double div = (fx_sv1 - fx_sv0) / a.fc.dx + (fy_sv0_p - fy_sv0) / a.fc.dy;
        const double pg_now = one_m_eps * cur.s_now.x * (mn_p.x - mn_m.x) / a.two_dy;
        const double pg_new = a.eps * s0 * (mw_p.x - mw_m.x) / a.two_dy;
        sv0 = cur.sv_now.x - a.dt * (div + pg_now + pg_new - tnd_sv.x);
      }
      {
        const double div = (fx_sv2 - fx_sv1) / a.fc.dx + (fy_sv1_p - fy_sv1) / a.fc.dy;
        const double pg_now = one_m_eps * cur.s_now.y * (mn_p.y - mn_m.y) / a.two_dy;
        const double pg_new = a.eps * s1 * (mw_p.y - mw_m.y) / a.two_dy;
        sv1 = cur.sv_now.y - a.dt * (div + pg_now + pg_new - tnd_sv.y);
      }
      if (!int0) { su0 = 0.0; sv0 = 0.0; }
      if (!int1) { su1 = 0.0; sv1 = 0.0; }
    }
    const double gam0 = gam_a.x, gam1 = gam_a.y;
    const bool touched = gam0 != 0.0 || gam1 != 0.0 || r_damp != 0.0;
    double2 s_ref = zero2, su_ref = zero2, sv_ref = zero2;
    if (touched && ld_own) {
      const double *pr = w_ref + (r & 1) * 64;
      s_ref = lds2(pr);
      su_ref = lds2(pr + REF_RING2 * 64);
      sv_ref = lds2(pr + 2 * REF_RING2 * 64);
    }
    if (ld_own && ((!int0 && gam0 != 1.0) || (!int1 && gam1 != 1.0))) {  // not reached with a Relaxed boundary
      const double2 ou = ldo2(a.su_new.p, o_c), ov = ldo2(a.sv_new.p, o_c);
      if (!int0 && gam0 != 1.0) { su0 = ou.x; sv0 = ov.x; }
      if (!int1 && gam1 != 1.0) { su1 = ou.y; sv1 = ov.y; }
    }
    if (gam0 != 0.0) {  // hb.enforce_raw, dycore.py:L686
      s0 = relax_point(gam0, s0, s_ref.x);
      su0 = relax_point(gam0, su0, su_ref.x);
      sv0 = relax_point(gam0, sv0, sv_ref.x);
    }
    if (gam1 != 0.0) {
      s1 = relax_point(gam1, s1, s_ref.y);
      su1 = relax_point(gam1, su1, su_ref.y);
      sv1 = relax_point(gam1, sv1, sv_ref.y);
    }
    if (r_damp != 0.0) {  // dycore.py:L694-L700
      s0 = damp_point(cur.s_now.x, s0, s_ref.x, r_damp, a.dt_full);
      su0 = damp_point(cur.su_now.x, su0, su_ref.x, r_damp, a.dt_full);
      sv0 = damp_point(cur.sv_now.x, sv0, sv_ref.x, r_damp, a.dt_full);
      s1 = damp_point(cur.s_now.y, s1, s_ref.y, r_damp, a.dt_full);
      su1 = damp_point(cur.su_now.y, su1, su_ref.y, r_damp, a.dt_full);
      sv1 = damp_point(cur.sv_now.y, sv1, sv_ref.y, r_damp, a.dt_full);
    }

    if (UVOUT) {
      // ---- velocity diagnosis (dwarfs/diagnostics.py:L219-L272) and stores
      const double su_l = __shfl_up_sync(0xffffffffu, su1, 1);
      const double s_l = __shfl_up_sync(0xffffffffu, s1, 1);
      if (out0 && r >= j0) {
        const int c1 = c0 + 1;
        const double u0 = c0 == 0 ? ldo(a.u_ref.p, o_c) : qdiv(su_l + su0, s_l + s0);
        double v0, v1;
        if (r == 0) {
          const double2 vr = ldo2(a.v_ref.p, o_c);
          v0 = vr.x; v1 = vr.y;
        } else {
          v0 = qdiv(sv_prev0 + sv0, s_prev0 + s0);
          v1 = qdiv(sv_prev1 + sv1, s_prev1 + s1);
        }
        if (out1) {
          const double u1 = qdiv(su0 + su1, s0 + s1);
          sto2(a.s_new.p, o_c, make_double2(s0, s1));
          sto2(a.su_new.p, o_c, make_double2(su0, su1));
          sto2(a.sv_new.p, o_c, make_double2(sv0, sv1));
          sto2(a.u_new.p, o_c, make_double2(u0, u1));
          sto2(a.v_new.p, o_c, make_double2(v0, v1));
          if (c1 == nx - 1) sto(a.u_new.p, o_c + 16u, ldo(a.u_ref.p, o_c + 16u));  // relaxed.py:L161-L175
          if (r == ny - 1) sto2(a.v_new.p, o_c + row, ldo2(a.v_ref.p, o_c + row));  // relaxed.py:L177-L191
        } else {  // c0 is the last column of an odd-sized row
          sto(a.s_new.p, o_c, s0);
          sto(a.su_new.p, o_c, su0);
          sto(a.sv_new.p, o_c, sv0);
          sto(a.u_new.p, o_c, u0);
          sto(a.v_new.p, o_c, v0);
          sto(a.u_new.p, o_c + 8u, ldo(a.u_ref.p, o_c + 8u));
          if (r == ny - 1) sto(a.v_new.p, o_c + row, ldo(a.v_ref.p, o_c + row));
        }
      }
      sv_prev0 = sv0; sv_prev1 = sv1;
      s_prev0 = s0; s_prev1 = s1;
    } else if (out0) {  // intermediate stage: s, su, sv only (r >= j0 always: no warm-up row)
      const bool put_s = !s_inplace || touched;
      if (out1) {
        if (put_s) sto2(a.s_new.p, o_c, make_double2(s0, s1));
        sto2(a.su_new.p, o_c, make_double2(su0, su1));
        sto2(a.sv_new.p, o_c, make_double2(sv0, sv1));
      } else {
        if (put_s) sto(a.s_new.p, o_c, s0);
        sto(a.su_new.p, o_c, su0);
        sto(a.sv_new.p, o_c, sv0);
      }
    }
    fy_su0 = fy_su0_p; fy_su1 = fy_su1_p;
    fy_sv0 = fy_sv0_p; fy_sv1 = fy_sv1_p;
    gam_a = gam_b; gam_b = gam_c;
    if (DERIVE) {
      si_a = si_b;
      si_b = cur.s_i2;
    }
    o_c += row;
  }
  cp_async_wait<0>();  // nothing may still be in flight into this block's shared memory
}

// TB200_MV_IMPL selects the momentum kernel: "window" (register windows), "ring" (warp-private
// shared-memory rings, one column per lane) or, by default, the two-columns-per-lane ring kernel
int mv_impl() {
  static int impl = -1;
  if (impl < 0) {
    const char *e = getenv("TB200_MV_IMPL");
    impl = e == nullptr ? 2 : strcmp(e, "window") == 0 ? 0 : strcmp(e, "ring") == 0 ? 1 : 2;
  }
  return impl;
}

// TB200_MV_BLOCK = 2x2 | 3x1 | 6x1: warps per block of the two-column kernel (side by side x
// strips).  Warps of one block start together and stay close in time, so the halo sectors two
// x-neighbours share are still in L2 when the second one asks; across blocks they are not.
// The alternatives to the default are instantiated for the benchmark configuration only
// (fifth-order fluxes, 64-row strips).
int mv_block() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("TB200_MV_BLOCK");
    // where the three shapes exist (fifth-order fluxes, 64-row strips: the benchmark configuration)
    // 2 x 2 is the measured optimum once the velocities are derived in the kernel (round 2, config 5:
    // 1.25 / 1.26 / 1.26 ms per stage against 1.24 / 1.36 / 1.36 for 3 x 1 and 1.26 / 1.37 / 1.36 for 6 x 1)
    v = e == nullptr ? 0 : strcmp(e, "2x2") == 0 ? 0 : strcmp(e, "3x1") == 0 ? 1 : strcmp(e, "6x1") == 0 ? 2 : 0;
  }
  return v;
}

// the two-column kernel moves 16-byte pairs: aligned bases, even pitches, one spare column
bool mv2_ok(const StageArgs &a) {
  const View *all[] = {&a.s_now, &a.su_now, &a.sv_now, &a.mtg_now, &a.su_int, &a.sv_int, &a.u_int,
                       &a.v_int, &a.s_new, &a.su_new, &a.sv_new, &a.u_new, &a.v_new, &a.s_ref,
                       &a.su_ref, &a.sv_ref, &a.u_ref, &a.v_ref, &a.mtg, &a.spre, &a.gamma};
  for (const View *v : all)
    if ((reinterpret_cast<uintptr_t>(v->p) & 15) != 0 || (v->s1 & 1) != 0 || (v->s2 & 1) != 0 ||
        v->s1 < a.nx + 1)
      return false;
  return true;
}

struct MvGeom {
  int impl, wx, wy, cols;  // kernel, warps per block along x / along y, owned columns per warp
};
template <int SCHEME, int LJ>
constexpr bool mv2_has_blocks() {  // are the non-default block shapes instantiated?
  return SCHEME == TB200_FLUX_FIFTH_ORDER_UPWIND && LJ == 64;
}
MvGeom mv_geom(const StageArgs &a, bool alt_blocks) {
  int impl = mv_impl();
  if (impl == 2 && !mv2_ok(a)) impl = 1;
  if (impl != 2) return MvGeom{impl, 4, 1, MV_COLS};
  const int b = (alt_blocks && !a.su_tnd.ok()) ? mv_block() : MV2_BLOCK_DEFAULT;
  return b == 0 ? MvGeom{2, 2, 2, MV2_COLS} : b == 1 ? MvGeom{2, 3, 1, MV2_COLS} : MvGeom{2, 6, 1, MV2_COLS};
}
// is the default path (kernels A + B + two-column MV) in charge?  Only it honours derive_uv_in /
// skip_uv_out; every other variant reads u_int / v_int and writes u_new / v_new like the
// reference, which is consistent as long as ALL stages of a run use the same variant (the
// selection depends on the process environment and on the storages' layout only).
bool lazy_uv_path(const StageArgs &a) {
  return s_impl() != 0 && a.nz <= 64 && stage_impl() == 0 && mv_impl() == 2 && mv2_ok(a);
}

template <int SCHEME, int LJ, int WX, int WY, bool DERIVE, bool UVOUT, bool TND = false>
int launch_mv2(const StageArgs &a, dim3 grid, cudaStream_t st) {
  const size_t smem = (size_t)WX * WY * WARP_DOUBLES2 * sizeof(double);
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(stage_mv2_kernel<SCHEME, LJ, WX, WY, DERIVE, UVOUT, TND>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("isentropic_stage_dry/MV2: %s", cudaGetErrorString(cudaGetLastError()));
      return TB200_ERR_CUDA;
    }
    configured = true;
  }
  stage_mv2_kernel<SCHEME, LJ, WX, WY, DERIVE, UVOUT, TND><<<grid, dim3(32 * WX * WY, 1, 1), smem, st>>>(a);
  return check_launch("isentropic_stage_dry/MV(2 columns)");
}
template <int SCHEME, int LJ, int WX, int WY>
int launch_mv2_flags(const StageArgs &a, dim3 grid, cudaStream_t st) {
  if constexpr (WX == MV2_DEF_WX && WY == MV2_DEF_WY) {  // slow tendencies: the default block shape only (mv_geom)
    if (a.su_tnd.ok()) {
      if (a.derive_uv)
        return a.skip_uv ? launch_mv2<SCHEME, LJ, WX, WY, true, false, true>(a, grid, st)
                         : launch_mv2<SCHEME, LJ, WX, WY, true, true, true>(a, grid, st);
      return a.skip_uv ? launch_mv2<SCHEME, LJ, WX, WY, false, false, true>(a, grid, st)
                       : launch_mv2<SCHEME, LJ, WX, WY, false, true, true>(a, grid, st);
    }
  }
  if (a.derive_uv)
    return a.skip_uv ? launch_mv2<SCHEME, LJ, WX, WY, true, false>(a, grid, st)
                     : launch_mv2<SCHEME, LJ, WX, WY, true, true>(a, grid, st);
  return a.skip_uv ? launch_mv2<SCHEME, LJ, WX, WY, false, false>(a, grid, st)
                   : launch_mv2<SCHEME, LJ, WX, WY, false, true>(a, grid, st);
}

// The momentum kernel over a rectangle of its block grid.
template <int SCHEME, int LJ>
int launch_mv_rect(StageArgs a, const MvGeom &g, int bx0, int bx1, int by0, int by1, cudaStream_t st) {
  if (bx1 <= bx0 || by1 <= by0) return TB200_OK;
  a.bx0 = bx0;
  a.by0 = by0;
  dim3 block(32 * g.wx * g.wy, 1, 1);
  dim3 grid(bx1 - bx0, by1 - by0, a.nz);
  if (g.impl == 2) {
    if constexpr (mv2_has_blocks<SCHEME, LJ>()) {
      if (g.wx == 2) return launch_mv2_flags<SCHEME, LJ, 2, 2>(a, grid, st);
      if (g.wx == 6) return launch_mv2_flags<SCHEME, LJ, 6, 1>(a, grid, st);
    }
    return launch_mv2_flags<SCHEME, LJ, MV2_DEF_WX, MV2_DEF_WY>(a, grid, st);
  }
  if (g.impl == 1) {
    stage_mv_ring_kernel<SCHEME, LJ><<<grid, block, g.wx * WARP_DOUBLES * sizeof(double), st>>>(a);
    return check_launch("isentropic_stage_dry/MV(ring)");
  }
  stage_mv_kernel<SCHEME, LJ><<<grid, block, 0, st>>>(a);
  return check_launch("isentropic_stage_dry/MV");
}

// part 0: the whole block grid; part 1: the blocks holding the a.rim columns / rows next to an
// edge with a neighbour (south and north strips over all columns, west and east blocks over the
// remaining strips); part 2: the interior rectangle.
template <int SCHEME, int LJ>
int launch_mv_part(const StageArgs &a, cudaStream_t st) {
  const MvGeom g = mv_geom(a, mv2_has_blocks<SCHEME, LJ>());
  const int cols = g.wx * g.cols, rows = g.wy * LJ;  // columns / rows per block
  const int gx = (a.nx + cols - 1) / cols, gy = (a.ny + rows - 1) / rows;
  if (a.part == 0) return launch_mv_rect<SCHEME, LJ>(a, g, 0, gx, 0, gy, st);
  // first interior block after the west rim / first east-rim block, likewise for the strips
  int xw = a.rim[0] > 0 ? (a.rim[0] + cols - 1) / cols : 0;
  int xe = a.rim[1] > 0 ? (a.nx - a.rim[1]) / cols : gx;
  int ys = a.rim[2] > 0 ? (a.rim[2] + rows - 1) / rows : 0;
  int yn = a.rim[3] > 0 ? (a.ny - a.rim[3]) / rows : gy;
  xe = max(min(xe, gx), xw);
  yn = max(min(yn, gy), ys);
  if (a.part == 2) return launch_mv_rect<SCHEME, LJ>(a, g, xw, xe, ys, yn, st);
  int rc = launch_mv_rect<SCHEME, LJ>(a, g, 0, gx, 0, ys, st);
  if (!rc) rc = launch_mv_rect<SCHEME, LJ>(a, g, 0, gx, yn, gy, st);
  if (!rc) rc = launch_mv_rect<SCHEME, LJ>(a, g, 0, xw, ys, yn, st);
  if (!rc) rc = launch_mv_rect<SCHEME, LJ>(a, g, xe, gx, ys, yn, st);
  return rc;
}

// Rows per warp strip (LJ) of the j-marching kernels A and MV.  64 is the measured optimum on
// large grids (1/64 of redundant warm-up rows; config 5), but a 161 x 161 x 60 grid (config 2)
// then yields only 1080 / 540 warps for the 148 SMs -- 1-2 warps per scheduler marching serially
// through 64 dependent rows, 5x slower per point than at config 5 (profiles/README.md, round 1d).
// Shorter strips trade warm-up redundancy for parallelism: the largest LJ of {64, 16, 8} that gives
// every scheduler TARGET warps is used.  TB200_LJ=64|16|8 forces one (experiments).
constexpr int SM_COUNT = 148, LJ_TARGET_WARPS = SM_COUNT * 4 * 4;
int lj_forced() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("TB200_LJ");
    v = e == nullptr ? 0 : atoi(e);
    if (v != 64 && v != 16 && v != 8) v = 0;
  }
  return v;
}
int pick_lj(long long strips_x, int ny, int nz) {
  if (lj_forced()) return lj_forced();
  for (int lj : {64, 16}) {
    if (strips_x * ((ny + lj - 1) / lj) * nz >= LJ_TARGET_WARPS) return lj;
  }
  return 8;
}

// TB200_B_IMPL=column|coop forces the scan kernel; by default the cooperative one runs when one
// thread per column would leave the schedulers with fewer than four warps each
bool b_coop(const StageArgs &a) {
  static int forced = -1;
  if (forced < 0) {
    const char *e = getenv("TB200_B_IMPL");
    forced = e == nullptr ? 0 : strcmp(e, "column") == 0 ? 1 : strcmp(e, "coop") == 0 ? 2 : 0;
  }
  if (forced) return forced == 2;
  return (long long)a.nx * a.ny < (long long)LJ_TARGET_WARPS * 32;
}

int a_impl() {  // TB200_A_IMPL=one|two (columns per lane; default one)
  static int impl = -1;
  if (impl < 0) {
    const char *e = getenv("TB200_A_IMPL");
    impl = (e != nullptr && strcmp(e, "two") == 0) ? 2 : 1;
  }
  return impl;
}

template <int SCHEME, int LJ>
int launch_a(const StageArgs &a, cudaStream_t st) {
  constexpr int WARPS = 4;
  if (a.s_tnd.ok()) {  // slow tendency of s: the one-column kernel
    const int chunks_t = (a.nx + A_COLS - 1) / A_COLS;
    dim3 block_t(32 * WARPS, 1, 1);
    dim3 grid_t((chunks_t + WARPS - 1) / WARPS, (a.ny + LJ - 1) / LJ, a.nz);
    if (a.derive_uv)
      stage_a_kernel<SCHEME, LJ, true, true><<<grid_t, block_t, 0, st>>>(a);
    else
      stage_a_kernel<SCHEME, LJ, false, true><<<grid_t, block_t, 0, st>>>(a);
    return check_launch("isentropic_stage_dry/A(tendency)");
  }
  if (a_impl() == 2 && a.a2_ok) {
    const int chunks2 = (a.nx + A2_COLS - 1) / A2_COLS;
    dim3 block2(32 * WARPS, 1, 1);
    dim3 grid2((chunks2 + WARPS - 1) / WARPS, (a.ny + LJ - 1) / LJ, a.nz);
    if (a.derive_uv)
      stage_a2_kernel<SCHEME, LJ, true><<<grid2, block2, 0, st>>>(a);
    else
      stage_a2_kernel<SCHEME, LJ, false><<<grid2, block2, 0, st>>>(a);
    return check_launch("isentropic_stage_dry/A(2 columns)");
  }
  const int chunks = (a.nx + A_COLS - 1) / A_COLS;
  dim3 block(32 * WARPS, 1, 1);
  dim3 grid((chunks + WARPS - 1) / WARPS, (a.ny + LJ - 1) / LJ, a.nz);
  if (a.derive_uv)
    stage_a_kernel<SCHEME, LJ, true><<<grid, block, 0, st>>>(a);
  else
    stage_a_kernel<SCHEME, LJ, false><<<grid, block, 0, st>>>(a);
  return check_launch("isentropic_stage_dry/A");
}

template <int SCHEME>
int launch_mv(const StageArgs &a, cudaStream_t st) {
  const MvGeom g = mv_geom(a, false);
  // the overlap parts of a decomposed run are laid out in blocks of 64-row strips; the earlier
  // momentum kernels (TB200_MV_IMPL=window|ring) are kept as they were measured, at 64 rows
  const int lj = (a.part != 0 || g.impl != 2) ? 64 : pick_lj((a.nx + g.cols - 1) / g.cols, a.ny, a.nz);
  if (lj == 16) return launch_mv_part<SCHEME, 16>(a, st);
  if (lj == 8) return launch_mv_part<SCHEME, 8>(a, st);
  return launch_mv_part<SCHEME, 64>(a, st);
}

template <int SCHEME>
int run_stage(const StageArgs &a, cudaStream_t st) {
  if (a.part == 2) return launch_mv<SCHEME>(a, st);  // the s-step and the scans ran with part 1
  prof_mark(0, st);
  if (s_impl() != 0 && a.nz <= 64) {
    {
      const int acols = (a_impl() == 2 && a.a2_ok && !a.s_tnd.ok()) ? A2_COLS : A_COLS;
      const int lj = pick_lj((a.nx + acols - 1) / acols, a.ny, a.nz);
      const int rc = lj == 16 ? launch_a<SCHEME, 16>(a, st)
                     : lj == 8 ? launch_a<SCHEME, 8>(a, st) : launch_a<SCHEME, 64>(a, st);
      if (rc) return rc;
      prof_mark(1, st);
    }
    if (a.ntr > 0) {  // the water constituents need the stage's s before kernel MV finalises it in place
      const int rc = launch_tracers<SCHEME>(a, st);
      if (rc) return rc;
    }
    if (a.periodic) {  // hb.enforce_field(s_new), rk3ws_si.py:L184-L189, for periodic.py:L98-L122
      tb200_field f;
      f.ptr = a.spre.p;
      f.shape[0] = a.spre.n0; f.shape[1] = a.spre.n1; f.shape[2] = a.nz;
      f.stride[0] = a.spre.s0; f.stride[1] = a.spre.s1; f.stride[2] = a.spre.s2;
      const int px = a.nx - 2 * a.nb, py = a.ny - 2 * a.nb;
      const int rc = tb200_periodic_enforce(&f, px, py, a.nb, px, py, st);
      if (rc) return rc;
    }
    if (b_coop(a)) {
      dim3 block(32, 8, 1);
      dim3 grid((a.nx + 31) / 32, a.ny, 1);
      if (a.nz <= 32)
        stage_b_coop_kernel<32><<<grid, block, 0, st>>>(a);
      else
        stage_b_coop_kernel<64><<<grid, block, 0, st>>>(a);
      int rc = check_launch("isentropic_stage_dry/B(coop)");
      if (rc) return rc;
    } else {
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
    const int rc = launch_mv<SCHEME>(a, st);
    prof_mark(3, st);
    g_prof.recorded = g_prof.on;
    return rc;
  }
}

bool covers(const View &v, int ni, int nj, int nk) {
  return v.ok() && v.n0 >= ni && v.n1 >= nj && v.n2 >= nk;
}

}  // namespace

// Do the kernels selected by the process environment honour derive_uv_in / skip_uv_out and the
// in-place s (scratch_s == s_new)?  1 on the default path (kernels A + B + two-column momentum
// kernel, columns of at most 64 layers), 0 when TB200_S_IMPL / TB200_MV_IMPL / TB200_STAGE_IMPL
// select an earlier variant -- a host then keeps the reference's data flow.
extern "C" int tb200_stage_lazy_velocities(int nz) {
  return s_impl() != 0 && nz <= 64 && stage_impl() == 0 && mv_impl() == 2 ? 1 : 0;
}

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

namespace {
int stage_entry(
    const tb200_isentropic_stage *cfg, const tb200_field *s_now, const tb200_field *su_now,
    const tb200_field *sv_now, const tb200_field *mtg_now, const tb200_field *s_int,
    const tb200_field *su_int, const tb200_field *sv_int, const tb200_field *u_int,
    const tb200_field *v_int, tb200_field *s_new, tb200_field *su_new, tb200_field *sv_new,
    tb200_field *u_new, tb200_field *v_new, const tb200_field *s_ref, const tb200_field *su_ref,
    const tb200_field *sv_ref, const tb200_field *u_ref, const tb200_field *v_ref,
    const tb200_field *gamma, const tb200_field *rmat, const tb200_field *hs,
    tb200_field *scratch_exn, tb200_field *scratch_mtg, tb200_field *scratch_s,
    const tb200_field *const *q_now, const tb200_field *const *q_int, tb200_field *const *q_new,
    const tb200_field *const *q_ref, void *stream) {
  TB200_REQUIRE(cfg != nullptr, "isentropic_stage_dry: NULL cfg");
  StageArgs a{};
  a.s_tnd = view(cfg->s_tnd); a.su_tnd = view(cfg->su_tnd); a.sv_tnd = view(cfg->sv_tnd);
  a.ntr = 0;
  if (q_now != nullptr) {
    TB200_REQUIRE(q_int != nullptr && q_new != nullptr && q_ref != nullptr,
                  "isentropic_stage_moist: NULL tracer array");
    a.ntr = 3;
    for (int t = 0; t < 3; ++t) {
      a.q_now[t] = view(q_now[t]); a.q_int[t] = view(q_int[t]);
      a.q_new[t] = view(q_new[t]); a.q_ref[t] = view(q_ref[t]);
    }
  }
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
  a.derive_uv = cfg->derive_uv_in != 0;
  a.skip_uv = cfg->skip_uv_out != 0;
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
  for (int t = 0; t < a.ntr; ++t) {
    const View *qs[] = {&a.q_now[t], &a.q_int[t], &a.q_new[t], &a.q_ref[t]};
    for (const View *v : qs) {
      TB200_REQUIRE(covers(*v, nx, ny, nz), "isentropic_stage_moist: a tracer field is NULL or too small");
      if (v->s0 != 1 || v->s1 != a.s_now.s1 || v->s2 != a.s_now.s2) {
        set_error("isentropic_stage_moist: the tracer fields must share the geometry of the other fields");
        return TB200_ERR_LAYOUT;
      }
    }
    TB200_REQUIRE(a.q_new[t].p != a.q_int[t].p && a.q_new[t].p != a.q_now[t].p,
                  "isentropic_stage_moist: q_new must not alias the stage inputs");
  }
  TB200_REQUIRE(a.ntr == 0 || (a.part == 0 && s_impl() != 0 && a.nz <= 64),
                "isentropic_stage_moist: needs the default kernel path (A + B), an unsplit stage and nz <= 64");
  a.a2_ok = mv2_ok(a) ? 1 : 0;
  a.periodic = cfg->periodic != 0;
  TB200_REQUIRE(!a.periodic || (a.part == 0 && lazy_uv_path(a) && cfg->skip_uv_out != 0 && !a.damp &&
                                nx >= 4 * a.nb && ny >= 4 * a.nb),
                "isentropic_stage_dry: a periodic stage needs part 0, the default kernel path, skip_uv_out, "
                "damp = 0 and nx, ny >= 4 nb");
  if (a.s_tnd.ok() || a.su_tnd.ok() || a.sv_tnd.ok()) {
    // slow tendencies (rk3ws_si.py:L105-L234 passes s_tnd, su_tnd, sv_tnd to K1 / K2): all three or
    // none, same geometry, dry stage, default kernel path
    const View *ts[] = {&a.s_tnd, &a.su_tnd, &a.sv_tnd};
    for (const View *v : ts) {
      TB200_REQUIRE(covers(*v, nx, ny, nz), "isentropic_stage_dry: s_tnd, su_tnd, sv_tnd must be given together "
                                            "and cover (nx, ny, nz)");
      if (v->s0 != 1 || v->s1 != a.s_now.s1 || v->s2 != a.s_now.s2 ||
          (reinterpret_cast<uintptr_t>(v->p) & 15) != 0) {
        set_error("isentropic_stage_dry: the tendency fields must share the geometry of the other fields");
        return TB200_ERR_LAYOUT;
      }
    }
    TB200_REQUIRE(a.ntr == 0 && a.part == 0 && lazy_uv_path(a),
                  "isentropic_stage_dry: slow tendencies need the dry stage, part 0 and the default kernel path");
  }
  if (!lazy_uv_path(a)) {
    // the other kernel variants keep the reference's data flow (see lazy_uv_path)
    TB200_REQUIRE(a.spre.p != a.s_new.p,
                  "isentropic_stage_dry: scratch_s may alias s_new only on the default kernel path");
    a.derive_uv = a.skip_uv = 0;
  }
  TB200_REQUIRE(a.spre.p != a.s_new.p || a.skip_uv,
                "isentropic_stage_dry: scratch_s may alias s_new only with skip_uv_out");
  TB200_REQUIRE(a.spre.p != a.s_new.p || a.part == 0,
                "isentropic_stage_dry: scratch_s may alias s_new only for an unsplit stage (part 0)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (cfg->flux_scheme) {
    case TB200_FLUX_UPWIND: return run_stage<TB200_FLUX_UPWIND>(a, st);
    case TB200_FLUX_CENTERED: return run_stage<TB200_FLUX_CENTERED>(a, st);
    case TB200_FLUX_THIRD_ORDER_UPWIND: return run_stage<TB200_FLUX_THIRD_ORDER_UPWIND>(a, st);
    default: return run_stage<TB200_FLUX_FIFTH_ORDER_UPWIND>(a, st);
  }
}
}  // namespace

extern "C" int tb200_isentropic_stage_dry(
    const tb200_isentropic_stage *cfg, const tb200_field *s_now, const tb200_field *su_now,
    const tb200_field *sv_now, const tb200_field *mtg_now, const tb200_field *s_int,
    const tb200_field *su_int, const tb200_field *sv_int, const tb200_field *u_int,
    const tb200_field *v_int, tb200_field *s_new, tb200_field *su_new, tb200_field *sv_new,
    tb200_field *u_new, tb200_field *v_new, const tb200_field *s_ref, const tb200_field *su_ref,
    const tb200_field *sv_ref, const tb200_field *u_ref, const tb200_field *v_ref,
    const tb200_field *gamma, const tb200_field *rmat, const tb200_field *hs,
    tb200_field *scratch_exn, tb200_field *scratch_mtg, tb200_field *scratch_s, void *stream) {
  return stage_entry(cfg, s_now, su_now, sv_now, mtg_now, s_int, su_int, sv_int, u_int, v_int, s_new,
                     su_new, sv_new, u_new, v_new, s_ref, su_ref, sv_ref, u_ref, v_ref, gamma, rmat, hs,
                     scratch_exn, scratch_mtg, scratch_s, nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int tb200_isentropic_stage_moist(
    const tb200_isentropic_stage *cfg, const tb200_field *s_now, const tb200_field *su_now,
    const tb200_field *sv_now, const tb200_field *mtg_now, const tb200_field *s_int,
    const tb200_field *su_int, const tb200_field *sv_int, const tb200_field *u_int,
    const tb200_field *v_int, tb200_field *s_new, tb200_field *su_new, tb200_field *sv_new,
    tb200_field *u_new, tb200_field *v_new, const tb200_field *s_ref, const tb200_field *su_ref,
    const tb200_field *sv_ref, const tb200_field *u_ref, const tb200_field *v_ref,
    const tb200_field *gamma, const tb200_field *rmat, const tb200_field *hs,
    tb200_field *scratch_exn, tb200_field *scratch_mtg, tb200_field *scratch_s,
    const tb200_field *const *q_now, const tb200_field *const *q_int, tb200_field *const *q_new,
    const tb200_field *const *q_ref, void *stream) {
  TB200_REQUIRE(q_now != nullptr && q_int != nullptr && q_new != nullptr && q_ref != nullptr,
                "isentropic_stage_moist: NULL tracer array");
  return stage_entry(cfg, s_now, su_now, sv_now, mtg_now, s_int, su_int, sv_int, u_int, v_int, s_new,
                     su_new, sv_new, u_new, v_new, s_ref, su_ref, sv_ref, u_ref, v_ref, gamma, rmat, hs,
                     scratch_exn, scratch_mtg, scratch_s, q_now, q_int, q_new, q_ref, stream);
}
