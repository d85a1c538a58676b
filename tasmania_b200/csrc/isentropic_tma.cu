// isentropic_tma.cu -- the fused dry RK stage as TMA + shared-memory row pipelines (sm_100a).
//
// Same arithmetic, same results (bitwise) as the register-window kernels of
// isentropic_fused.cu; different data movement.  Those kernels spend one integer
// address computation per global load -- 36 % of their issued instructions -- and rely on
// register double-buffering for latency hiding.  Here the fields with spatial reuse are
// staged by the TMA unit:
//
//   * a CTA (4 warps) owns a strip of 120 columns of one level and marches along j;
//   * one elected thread issues cp.async.bulk.tensor (3-D tensor maps over the storages, box
//     128 columns x 2 rows x 1 level = 2 KB per field) into a ring of NG row groups in
//     shared memory, completion signalled on one mbarrier per group; out-of-range columns /
//     rows are zero-filled by the TMA unit, so there is no edge clamping in the kernel;
//   * the warps read neighbours from the ring with immediate-offset LDS (the ring geometry is a
//     compile-time constant: no address arithmetic), keep the y-stencil in register windows
//     (one new row per step), evaluate every face flux once (x faces shared by shuffle, y faces
//     carried to the next row) and store the results with coalesced STG;
//   * fields without reuse (s_pre, s_now, su_now, sv_now, the window's leading row) are plain
//     coalesced loads issued one row ahead.
//
// Stage = kernel S (isentropic_fused.cu: s-step, relaxation, column scans) + kernel C below
// (momentum step, relaxation, damping, velocities).
#include <cuda.h>

#include "stage.cuh"

using namespace tb200;

namespace {

constexpr int TW = 128;    // tile width [columns]: one shared-memory row = 1 KB
constexpr int HXL = 4;     // the tile starts 4 columns left of the CTA's first owned column
constexpr int WARPS = 4;
constexpr int WCOLS = 30;  // owned columns per warp (lanes 1..30; lanes 0 and 31 are helpers)
constexpr int RG = 2;      // rows per TMA group
constexpr int R_SU = 16;   // ring rows of su_int / sv_int (y stencil: rows r-E+1 .. r+E)
constexpr int R_MN = 8;    // ring rows of mtg_now / mtg_new (rows r, r+1)
constexpr int LEAD = 4;    // the rings start LEAD rows below the first computed row
constexpr int LJ = 64;     // rows per CTA (plus one warm-up row)
constexpr int CTAS_PER_SM = 4;
constexpr int PF_ROWS = 4;  // L2 prefetch distance of the direct (own-column) loads [rows]
constexpr int NMAPS_C = 4;  // su_int, sv_int, mtg_now, mtg_new

struct TmaMaps {
  CUtensorMap m[NMAPS_C];
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ unsigned smem_u32(const void *p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 3-D tiled bulk tensor load global -> shared, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// own-column values of one row (no reuse: plain coalesced loads, one row ahead of their use)
struct OwnColumn {
  double u, v, gam, s_pre, s_now, su_now, sv_now;
  double s_ref, su_ref, sv_ref;  // only loaded in the damping layer (CTA-uniform)
};

// ---------------------------------------------------------------- kernel C
// Momentum step + second relaxation + Rayleigh damping + velocity diagnosis (see the header).
// Shared memory: su_int and sv_int in rings of R_SU rows, mtg_now and mtg_new in rings of R_MN
// rows, all 128 columns wide, filled by TMA in groups of two rows; ring row of grid row rho =
// (rho - base) mod R with base = r0 - LEAD, so a stencil row is one AND + one shift away.
// No register windows and (almost) no loop-carried state: the rotating values live in the
// rings, which is what removes the register-move and address-arithmetic overhead of the
// register-window kernel.
template <int SCHEME>
__global__ void __launch_bounds__(WARPS * 32, CTAS_PER_SM)
    stage_c_kernel(const StageArgs a, const __grid_constant__ TmaMaps maps) {
  using F = Flux<SCHEME>;
  constexpr int E = F::extent;
  constexpr int NW = 2 * E;
  constexpr int NG_SU = R_SU / RG, NG_MN = R_MN / RG;
  constexpr unsigned SU_GROUP_BYTES = 2 * RG * TW * 8, MN_GROUP_BYTES = 2 * RG * TW * 8;
  static_assert(E <= LEAD - 1 && RG == 2, "ring geometry");

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double *su_ring = reinterpret_cast<double *>(smem_raw);  // [R_SU][TW]
  double *sv_ring = su_ring + R_SU * TW;
  double *mn_ring = sv_ring + R_SU * TW;                   // [R_MN][TW]
  double *mw_ring = mn_ring + R_MN * TW;
  uint64_t *full_su = reinterpret_cast<uint64_t *>(mw_ring + R_MN * TW);
  uint64_t *full_mn = full_su + NG_SU;
  uint64_t *empty = full_mn + NG_MN;  // [NG_MN]: one release per processed group

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int c0 = blockIdx.x * (WARPS * WCOLS);  // first owned column of the CTA
  const int c = c0 + w * WCOLS - 1 + lane;      // this lane's column (lane 0: left helper)
  const int tx = w * WCOLS + HXL - 1 + lane;    // its position in the tile
  const int j0 = blockIdx.y * LJ;
  const int jend = min(j0 + LJ, a.ny);
  const int k = blockIdx.z;
  const int nx = a.nx, ny = a.ny, nb = a.nb;
  const int r0 = j0 > 0 ? j0 - 1 : 0;  // first row computed (warm-up row unless j0 == 0)
  const int base = r0 - LEAD;          // grid row of ring row 0 (may be negative: TMA zero-fills)

  // ---- producer: group h of either ring = grid rows base + 2h, base + 2h + 1
  const int last_su = (jend - 1 + E - base) >> 1;  // rows up to jend-1+E are read
  const int last_mn = (jend - base) >> 1;          // rows up to jend
  constexpr int FIRST_MN = LEAD / RG;              // mtg rows below r0 are never read from the ring
  auto issue_su = [&](int h) {
    uint64_t *bar = &full_su[h % NG_SU];
    const int slot = (h % NG_SU) * RG * TW;
    mbar_expect_tx(bar, SU_GROUP_BYTES);
    tma_load_3d(su_ring + slot, &maps.m[0], c0 - HXL, base + h * RG, k, bar);
    tma_load_3d(sv_ring + slot, &maps.m[1], c0 - HXL, base + h * RG, k, bar);
  };
  auto issue_mn = [&](int h) {
    uint64_t *bar = &full_mn[h % NG_MN];
    const int slot = (h % NG_MN) * RG * TW;
    mbar_expect_tx(bar, MN_GROUP_BYTES);
    tma_load_3d(mn_ring + slot, &maps.m[2], c0 - HXL, base + h * RG, k, bar);
    tma_load_3d(mw_ring + slot, &maps.m[3], c0 - HXL, base + h * RG, k, bar);
  };
  if (tid == 0) {
#pragma unroll
    for (int g = 0; g < NG_SU; ++g) mbar_init(&full_su[g], 1);
#pragma unroll
    for (int g = 0; g < NG_MN; ++g) {
      mbar_init(&full_mn[g], 1);
      mbar_init(&empty[g], WARPS);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int h = 0; h < NG_SU && h <= last_su; ++h) issue_su(h);
    for (int h = FIRST_MN; h < FIRST_MN + NG_MN && h <= last_mn; ++h) issue_mn(h);
  }

  const bool out_lane = lane >= 1 && lane <= WCOLS && c < nx;
  const bool col_int = c >= nb && c < nx - nb;
  const int cc = min(max(c, 0), nx - 1);  // own column clamped into the row (direct loads)

  const unsigned row = (unsigned)a.s_now.s1 * 8u;  // bytes per row (all 3-D fields)
  const unsigned plane = (unsigned)k * (unsigned)a.s_now.s2 * 8u;
  const unsigned grow = (unsigned)a.gamma.s1 * 8u;
  const double r_damp = a.damp ? a.rmat.ld(0, 0, k) : 0.0;
  const bool damp_level = r_damp != 0.0;  // CTA-uniform
  const double one_m_eps = 1.0 - a.eps;

  unsigned o_c = plane + (unsigned)r0 * row + (unsigned)cc * 8u;  // own column, row r
  unsigned o_g = (unsigned)r0 * grow + (unsigned)cc * 8u;         // gamma (2-D)

  int landed_su = 0, landed_mn = FIRST_MN;  // groups whose data has arrived
  auto wait_su = [&](int h) {
    while (landed_su <= h) {
      mbar_wait(&full_su[landed_su % NG_SU], (landed_su / NG_SU) & 1);
      ++landed_su;
    }
  };
  auto wait_mn = [&](int h) {
    while (landed_mn <= h) {
      mbar_wait(&full_mn[landed_mn % NG_MN], ((landed_mn - FIRST_MN) / NG_MN) & 1);
      ++landed_mn;
    }
  };
  // pointers to this lane's entry of ring row (rho - base)
  auto su_row = [&](int q) { return su_ring + (q & (R_SU - 1)) * TW + tx; };
  auto mn_row = [&](int q) { return mn_ring + (q & (R_MN - 1)) * TW + tx; };
  constexpr int SV_OFF = R_SU * TW, MW_OFF = R_MN * TW;  // sv / mtg_new relative to su / mtg_now

  // ---- prologue: y-face flux of row r0 (rows r0-E .. r0+E-1) and the Montgomery row r0-1
  wait_su((LEAD + E - 1) >> 1);
  double fy_su, fy_sv, mn_m, mw_m;
  {
    double ysu[NW], ysv[NW];
#pragma unroll
    for (int m = 0; m < NW; ++m) {
      const double *p = su_row(LEAD - E + m);
      ysu[m] = p[0];
      ysv[m] = p[SV_OFF];
    }
    const double vq = F::prep(ldo(a.v_int.p, o_c), a.fc);
    fy_su = F::eval_v(vq, ysu);
    fy_sv = F::eval_v(vq, ysv);
    const unsigned o_m = plane + (unsigned)max(r0 - 1, 0) * row + (unsigned)cc * 8u;
    mn_m = ldo(a.mtg_now.p, o_m);
    mw_m = ldo(a.mtg.p, o_m);
  }
  double sv_prev = 0.0, s_prev = 0.0;
  int next_su = NG_SU, next_mn = FIRST_MN + NG_MN;  // next groups to issue (thread 0)

  // DRAM -> L2 prefetch of the own-column rows PF_ROWS ahead, by warp 0 only: ten 128-byte lines
  // cover the CTA's columns of one field; lanes 0-9 / 10-19 / 20-29 take one field each
  const double *pf_a = nullptr, *pf_b = nullptr, *pf_c = nullptr;
  if (w == 0 && lane < 30) {
    const int f = lane / 10;
    pf_a = f == 0 ? a.spre.p : f == 1 ? a.s_now.p : a.su_now.p;
    pf_b = f == 0 ? a.sv_now.p : f == 1 ? a.u_int.p : a.v_int.p;
    pf_c = !damp_level ? nullptr : f == 0 ? a.s_ref.p : f == 1 ? a.su_ref.p : a.sv_ref.p;
  }
  unsigned o_pf = plane + (unsigned)(r0 + PF_ROWS) * row +
                  (unsigned)((max(c0 - 1, 0) & ~15) * 8 + (lane % 10) * 128);

  auto load_own = [&](unsigned oc, unsigned og) {
    OwnColumn L;
    L.u = ldo(a.u_int.p, oc);
    L.v = ldo(a.v_int.p, oc + row);
    L.gam = ldo(a.gamma.p, og);
    L.s_pre = ldo(a.spre.p, oc);
    L.s_now = ldo(a.s_now.p, oc);
    L.su_now = ldo(a.su_now.p, oc);
    L.sv_now = ldo(a.sv_now.p, oc);
    L.s_ref = L.su_ref = L.sv_ref = 0.0;
    if (damp_level) {
      L.s_ref = ldo(a.s_ref.p, oc);
      L.su_ref = ldo(a.su_ref.p, oc);
      L.sv_ref = ldo(a.sv_ref.p, oc);
    }
    return L;
  };
  OwnColumn nxt = load_own(o_c, o_g);
  for (int r = r0; r < jend; ++r) {
    const int q = r - base;  // ring row of grid row r (before wrapping)
    // ---- own-column loads (L2 hits thanks to the prefetch), requested one row ahead
    const OwnColumn cur = nxt;
    nxt = load_own(o_c + row, o_g + grow);
    const double u_c = cur.u, v_n = cur.v, gam = cur.gam, s_pre = cur.s_pre, s_now = cur.s_now,
                 su_now = cur.su_now, sv_now = cur.sv_now;
    double s_ref = cur.s_ref, su_ref = cur.su_ref, sv_ref = cur.sv_ref;
    if (!damp_level && gam != 0.0) {  // relaxation band below the damping layer
      s_ref = ldo(a.s_ref.p, o_c);
      su_ref = ldo(a.su_ref.p, o_c);
      sv_ref = ldo(a.sv_ref.p, o_c);
    }
    if (w == 0) {
      if (pf_a != nullptr) prefetch_l2(pf_a, o_pf);
      if (pf_b != nullptr) prefetch_l2(pf_b, o_pf);
      if (pf_c != nullptr) prefetch_l2(pf_c, o_pf);
      o_pf += row;
    }
    wait_su((q + E) >> 1);  // rows up to r+E
    wait_mn((q + 1) >> 1);  // rows r, r+1

    // ---- y-face r+1: rows r-E+1 .. r+E of this column, straight from the ring
    double ysu[NW], ysv[NW];
#pragma unroll
    for (int m = 0; m < NW; ++m) {
      const double *p = su_row(q - E + 1 + m);
      ysu[m] = p[0];
      ysv[m] = p[SV_OFF];
    }
    const double vq = F::prep(v_n, a.fc);
    const double fy_su_p = F::eval_v(vq, ysu);
    const double fy_sv_p = F::eval_v(vq, ysv);

    // ---- left x-face of column c at row r: phi[c-E .. c+E-1]; phi[c] = ysu[E-1]
    const double uq = F::prep(u_c, a.fc);
    double xs[NW], ys[NW];
    {
      const double *p = su_row(q);
#pragma unroll
      for (int m = 0; m < NW; ++m) {
        xs[m] = m == E ? ysu[E - 1] : p[m - E];
        ys[m] = m == E ? ysv[E - 1] : p[SV_OFF + m - E];
      }
    }
    const double fx_su = F::eval_v(uq, xs);
    const double fx_sv = F::eval_v(uq, ys);
    const double fx_su_p = __shfl_down_sync(0xffffffffu, fx_su, 1);
    const double fx_sv_p = __shfl_down_sync(0xffffffffu, fx_sv, 1);

    // ---- point update (prognostics/utils.py:L191-L204)
    const double *pm = mn_row(q), *pmp = mn_row(q + 1);
    const double mn_p = pmp[0], mw_p = pmp[MW_OFF];
    const bool interior = col_int && r >= nb && r < ny - nb;
    double s = s_pre, su = 0.0, sv = 0.0;
    if (interior) {
      {
        const double div = (fx_su_p - fx_su) / a.fc.dx + (fy_su_p - fy_su) / a.fc.dy;
        const double pg_now = one_m_eps * s_now * (pm[1] - pm[-1]) / a.two_dx;
        const double pg_new = a.eps * s * (pm[MW_OFF + 1] - pm[MW_OFF - 1]) / a.two_dx;
        su = su_now - a.dt * (div + pg_now + pg_new - 0.0);
      }
      {
        const double div = (fx_sv_p - fx_sv) / a.fc.dx + (fy_sv_p - fy_sv) / a.fc.dy;
        const double pg_now = one_m_eps * s_now * (mn_p - mn_m) / a.two_dy;
        const double pg_new = a.eps * s * (mw_p - mw_m) / a.two_dy;
        sv = sv_now - a.dt * (div + pg_now + pg_new - 0.0);
      }
    }
    mn_m = pm[0];  // row r becomes row r-1 of the next iteration
    mw_m = pm[MW_OFF];
    if (!interior && gam != 1.0) {  // not reached with a Relaxed boundary (gamma == 1 there)
      su = ldo(a.su_new.p, o_c);
      sv = ldo(a.sv_new.p, o_c);
    }
    if (gam != 0.0) {  // hb.enforce_raw, dycore.py:L686
      s = relax_point(gam, s, s_ref);
      su = relax_point(gam, su, su_ref);
      sv = relax_point(gam, sv, sv_ref);
    }
    if (damp_level) {  // dycore.py:L694-L700
      s = damp_point(s_now, s, s_ref, r_damp, a.dt_full);
      su = damp_point(su_now, su, su_ref, r_damp, a.dt_full);
      sv = damp_point(sv_now, sv, sv_ref, r_damp, a.dt_full);
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
    o_c += row; o_g += grow;

    // ---- after the second row of a group: every warp releases it; thread 0 refills the
    // freed ring slots as soon as all four warps have (no block-wide barrier)
    if (q & 1) {
      const int P = q >> 1, n = P - FIRST_MN;  // processed group, counted from 0
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[n % NG_MN]);
      if (tid == 0) {
        // dead after this group: su/sv groups <= P-1 (their rows are below r-E+2 for every
        // scheme), mtg groups <= P; refill the freed slots with the next groups in line
        const bool more = (next_su <= last_su && next_su - NG_SU <= P - 1) ||
                          (next_mn <= last_mn && next_mn - NG_MN <= P);
        if (more) mbar_wait(&empty[n % NG_MN], (n / NG_MN) & 1);
        while (next_su <= last_su && next_su - NG_SU <= P - 1) issue_su(next_su++);
        while (next_mn <= last_mn && next_mn - NG_MN <= P) issue_mn(next_mn++);
      }
    }
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 3-D map over a storage: dims (row pitch, rows per plane, planes), box (TW, RG, 1)
bool make_map(CUtensorMap *map, const View &v) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)v.s1, (cuuint64_t)(v.s2 / v.s1), (cuuint64_t)v.n2};
  const cuuint64_t strides[2] = {(cuuint64_t)v.s1 * 8, (cuuint64_t)v.s2 * 8};
  const cuuint32_t box[3] = {TW, RG, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, v.p, dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int SCHEME>
int launch_c(const StageArgs &a, cudaStream_t st) {
  TmaMaps maps;
  const View *fields[NMAPS_C] = {&a.su_int, &a.sv_int, &a.mtg_now, &a.mtg};
  for (int f = 0; f < NMAPS_C; ++f)
    if (!make_map(&maps.m[f], *fields[f])) return -1;
  constexpr size_t smem = (size_t)(2 * R_SU + 2 * R_MN) * TW * 8 + (R_SU / RG + 2 * (R_MN / RG)) * sizeof(uint64_t);
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(stage_c_kernel<SCHEME>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess) {
      cudaGetLastError();
      return -1;
    }
    configured = true;
  }
  dim3 block(WARPS * 32, 1, 1);
  dim3 grid((a.nx + WARPS * WCOLS - 1) / (WARPS * WCOLS), (a.ny + LJ - 1) / LJ, a.nz);
  stage_c_kernel<SCHEME><<<grid, block, smem, st>>>(a, maps);
  return check_launch("isentropic_stage_dry/C");
}

}  // namespace

namespace tb200 {

int launch_stage_c(const StageArgs &a, int scheme, cudaStream_t st) {
  // the tensor maps need 16-byte aligned bases and row / plane pitches
  const View *fields[] = {&a.su_int, &a.sv_int, &a.mtg_now, &a.mtg};
  for (const View *v : fields)
    if ((reinterpret_cast<uintptr_t>(v->p) & 15) != 0 || (v->s1 & 1) != 0 || (v->s2 & 1) != 0) return -1;
  switch (scheme) {
    case TB200_FLUX_UPWIND: return launch_c<TB200_FLUX_UPWIND>(a, st);
    case TB200_FLUX_CENTERED: return launch_c<TB200_FLUX_CENTERED>(a, st);
    case TB200_FLUX_THIRD_ORDER_UPWIND: return launch_c<TB200_FLUX_THIRD_ORDER_UPWIND>(a, st);
    default: return launch_c<TB200_FLUX_FIFTH_ORDER_UPWIND>(a, st);
  }
}

}  // namespace tb200
