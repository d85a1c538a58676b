// horizontal.cu -- K8 horizontal diffusion (2nd / 4th order) and K9 horizontal smoothing
// (1st..3rd order): cross-shaped (i, j) stencils, no coupling in k.
//
// Roofline: HBM.  Algorithmic bytes 16 B/point (read phi once, write the tendency once;
// gamma is rank-1 in k by construction).  The tiled kernel stages a
// (TX + 2H) x (TY + 2H) tile of phi in shared memory per k-level so every phi value is
// fetched from L2/HBM once per CTA; threads keep i as the fast axis (coalesced 128-byte
// rows), gamma comes through the read-only path.  Grid = (tiles_i, tiles_j, nk) is far
// larger than 148 SMs x resident CTAs for the sizes of interest, so tail effects vanish.
#include <stdlib.h>
#include <string.h>

#include "stencil_math.cuh"

using namespace tb200;

namespace {

// ---- point formulas ------------------------------------------------------------------
// second_order.py:L101-L104
__device__ __forceinline__ double lap2(double g, double c, double im1, double ip1, double jm1,
                                       double jp1, const CDiv &dx, const CDiv &dy) {
  // dx, dy hold the full denominators dx*dx, dy*dy
  return g * ((im1 - 2.0 * c + ip1) / dx + (jm1 - 2.0 * c + jp1) / dy);
}
// fourth_order.py:L104-L122
__device__ __forceinline__ double lap4(double g, double c, double im2, double im1, double ip1,
                                       double ip2, double jm2, double jm1, double jp1,
                                       double jp2, const CDiv &dx, const CDiv &dy) {
  // dx, dy hold the full denominators 12*dx*dx, 12*dy*dy
  return g * ((-im2 + 16.0 * im1 - 30.0 * c + 16.0 * ip1 - ip2) / dx +
              (-jm2 + 16.0 * jm1 - 30.0 * c + 16.0 * jp1 - jp2) / dy);
}

constexpr int TX = 64, TY = 8;

// Shared-memory tiled cross stencil.  OP: 2/4 = diffusion order, 11/12/13 = smoothing 1..3.
template <int OP>
struct Halo {
  static constexpr int value = OP == 2 ? 1 : OP == 4 ? 2 : OP - 10;
};

template <int OP>
__global__ void __launch_bounds__(TX *TY)
    cross_kernel(View phi, View gam, View out, CDiv dx, CDiv dy, int overwrite, int rim,
                 int i0, int j0, int k0, int di, int dj, int dk, int ri, int rj) {
  constexpr int H = Halo<OP>::value;
  constexpr int SX = TX + 2 * H, SY = TY + 2 * H;
  __shared__ double tile[SY][SX + 1];

  // with rim copy the launch covers the whole (ri, rj) box, otherwise [i0, i0+di) x ...
  const int bi = rim ? 0 : i0, bj = rim ? 0 : j0;
  const int ti = bi + blockIdx.x * TX, tj = bj + blockIdx.y * TY;  // tile origin
  const int lim_i = rim ? ri : i0 + di, lim_j = rim ? rj : j0 + dj;
  const int tid = threadIdx.y * TX + threadIdx.x;

  for (int k = k0 + blockIdx.z; k < k0 + dk; k += gridDim.z) {
    // cooperative, row-coalesced load of the tile + halo (clamped to the storage)
    for (int t = tid; t < SX * SY; t += TX * TY) {
      const int ly = t / SX, lx = t - ly * SX;
      const int gi = ti + lx - H, gj = tj + ly - H;
      double v = 0.0;
      if (gi >= 0 && gi < phi.n0 && gj >= 0 && gj < phi.n1) v = phi.ld(gi, gj, k);
      tile[ly][lx] = v;
    }
    __syncthreads();

    const int i = ti + threadIdx.x, j = tj + threadIdx.y;
    if (i < lim_i && j < lim_j) {
      const int x = threadIdx.x + H, y = threadIdx.y + H;
      const double c = tile[y][x];
      const bool inside = i >= i0 && i < i0 + di && j >= j0 && j < j0 + dj;
      if (inside) {
        const double g = gam.ld(i, j, k);
        double r;
        if (OP == 2) {
          r = lap2(g, c, tile[y][x - 1], tile[y][x + 1], tile[y - 1][x], tile[y + 1][x], dx, dy);
        } else if (OP == 4) {
          r = lap4(g, c, tile[y][x - 2], tile[y][x - 1], tile[y][x + 1], tile[y][x + 2],
                   tile[y - 2][x], tile[y - 1][x], tile[y + 1][x], tile[y + 2][x], dx, dy);
        } else if (OP == 11) {  // first_order.py:L124-L126
          r = (1.0 - g) * c +
              0.25 * g * (tile[y][x - 1] + tile[y][x + 1] + tile[y - 1][x] + tile[y + 1][x]);
        } else if (OP == 12) {  // second_order.py:L126-L139
          r = (1.0 - 0.75 * g) * c +
              0.0625 * g *
                  (-tile[y][x - 2] + 4.0 * tile[y][x - 1] - tile[y][x + 2] +
                   4.0 * tile[y][x + 1] - tile[y - 2][x] + 4.0 * tile[y - 1][x] -
                   tile[y + 2][x] + 4.0 * tile[y + 1][x]);
        } else {  // third_order.py:L133-L150
          r = (1.0 - 0.625 * g) * c +
              0.015625 * g *
                  (tile[y][x - 3] - 6.0 * tile[y][x - 2] + 15.0 * tile[y][x - 1] +
                   tile[y][x + 3] - 6.0 * tile[y][x + 2] + 15.0 * tile[y][x + 1] +
                   tile[y - 3][x] - 6.0 * tile[y - 2][x] + 15.0 * tile[y - 1][x] +
                   tile[y + 3][x] - 6.0 * tile[y + 2][x] + 15.0 * tile[y + 1][x]);
        }
        if ((OP == 2 || OP == 4) && !overwrite) r = out(i, j, k) + r;  // generics.py:L38-L40
        out(i, j, k) = r;
      } else if (rim) {
        out(i, j, k) = c;  // the four `copy` launches of HorizontalSmoothing.__call__
      }
    }
    __syncthreads();
  }
}

// ---- the class-less `diffusion` stencil of the reference (a hyperdiffusion filter:
// framework/subclasses/stencil_definitions/diffusion.py:L31-L55): laplacian of the laplacian,
// its x / y differences as fluxes, phi + alpha * divergence of the fluxes.  A radius-3 diamond
// per point; not used by any component of the benchmark models, so it is written for clarity
// (every intermediate recomputed per thread from L1-resident loads, in the reference's order of
// operations), not for speed.
struct HyperOp {
  View phi, out;
  double alpha;
  int i0, j0, k0;
  __device__ __forceinline__ double lap(int i, int j, int k) const {
    return -4.0 * phi.ld(i, j, k) + phi.ld(i - 1, j, k) + phi.ld(i + 1, j, k) + phi.ld(i, j - 1, k) +
           phi.ld(i, j + 1, k);
  }
  __device__ __forceinline__ double bilap(int i, int j, int k) const {
    return -4.0 * lap(i, j, k) + lap(i - 1, j, k) + lap(i + 1, j, k) + lap(i, j - 1, k) + lap(i, j + 1, k);
  }
  __device__ void operator()(int ri, int rj, int rk) const {
    const int i = i0 + ri, j = j0 + rj, k = k0 + rk;
    const double bc = bilap(i, j, k);
    const double fx_hi = bilap(i + 1, j, k) - bc, fx_lo = bc - bilap(i - 1, j, k);
    const double fy_hi = bilap(i, j + 1, k) - bc, fy_lo = bc - bilap(i, j - 1, k);
    out(i, j, k) = phi.ld(i, j, k) + alpha * (fx_hi - fx_lo + fy_hi - fy_lo);
  }
};

// ---- diffusion, marching variant (TB200_DIFF_IMPL=march; 4.2 ms at 4096 x 4096 x 64, 0.62 of HBM).
// One thread per column i, marching along j over a strip of LJ rows with the 2H+1 rows of its
// own column in registers: every phi value is requested from L2/HBM once per strip (+ 2H halo
// rows per strip, 6 % at LJ = 64) and once more per x-neighbour from L1 (the neighbouring lanes'
// lines of the same row); no shared memory, no barrier, one running pointer per array.  Same
// point formulas, hence the same bits, as the tiled kernel.
template <int OP, int LJ>
__global__ void __launch_bounds__(128)
    march_kernel(View phi, View gam, View out, CDiv dx, CDiv dy, int overwrite, int i0, int j0,
                 int k0, int di, int dj, int dk) {
  constexpr int H = Halo<OP>::value;
  static_assert(OP == 2 || OP == 4, "diffusion only");
  const int i = i0 + blockIdx.x * 128 + threadIdx.x;
  if (i >= i0 + di) return;  // no shuffles and no barriers below
  const int js = j0 + blockIdx.y * LJ, je = min(js + LJ, j0 + dj);
  const long long s0 = phi.s0, s1 = phi.s1;
  for (int k = k0 + blockIdx.z; k < k0 + dk; k += gridDim.z) {
    const double *pr = phi.p + (i * s0 + js * s1 + k * phi.s2);  // (i, j, k), j running
    const double *pg = gam.p + (i * gam.s0 + js * gam.s1 + k * gam.s2);
    double *po = out.p + (i * out.s0 + js * out.s1 + k * out.s2);
    double w[2 * H + 1];  // rows j-H .. j+H of the own column
#pragma unroll
    for (int m = 0; m < 2 * H; ++m) w[m + 1] = __ldg(pr + (m - H) * s1);
    for (int j = js; j < je; ++j) {
#pragma unroll
      for (int m = 0; m < 2 * H; ++m) w[m] = w[m + 1];
      w[2 * H] = __ldg(pr + H * s1);
      const double g = __ldg(pg);
      double r;
      if (OP == 2) {
        r = lap2(g, w[H], __ldg(pr - s0), __ldg(pr + s0), w[H - 1], w[H + 1], dx, dy);
      } else {
        r = lap4(g, w[H], __ldg(pr - 2 * s0), __ldg(pr - s0), __ldg(pr + s0), __ldg(pr + 2 * s0),
                 w[H - 2], w[H - 1], w[H + 1], w[H + 2], dx, dy);
      }
      if (!overwrite) r = *po + r;  // generics.py:L38-L40
      *po = r;
      pr += s1;
      pg += gam.s1;
      po += out.s1;
    }
  }
}

// ---- diffusion, two columns per lane (default whenever the layout allows it).
// The one-column marching kernel above leaves the memory system idle: its only first-touch load
// per row is consumed in the same iteration, so a warp has 256 bytes in flight and the launch
// sits at 4.2 TB/s / 61 % warps active (profiles/r02_c4_ncu.md).  Here
//   * a lane owns TWO adjacent columns (an aligned pair): every access is LDG.128 / STG.128;
//   * the new row of the own pair (row j + H + 1) is requested one iteration before it enters
//     the window, and its line is pulled DRAM -> L2 PF2 rows earlier (prefetch.global.L2), so
//     the window load sees L2 latency and several rows per warp are in flight;
//   * the x-neighbours are the pairs of the lanes to the left and right: two more LDG.128 on
//     lines that entered L1 H + 1 rows ago;
//   * a block is eight warps SIDE BY SIDE (512 columns = 4 KB contiguous per row), no shared
//     memory, no barrier, no shuffle.
// Every phi value leaves HBM once per strip (+ 2H halo rows per LJ-row strip).  Same point
// formulas as the other kernels, hence the same bits.
// Requirements (checked on the host, else the one-column kernels run): unit i-stride, 16-byte
// aligned bases, even row / plane pitches of phi and out.
constexpr int M2_WARPS = 8, M2_PF = 4;

__device__ __forceinline__ double2 ldg2(const double *p) {
  return __ldg(reinterpret_cast<const double2 *>(p));
}

template <int OP, int LJ>
__global__ void __launch_bounds__(32 * M2_WARPS, 4)
    march2_kernel(View phi, View gam, View out, CDiv dx, CDiv dy, int overwrite, int i0, int j0,
                  int k0, int di, int dj, int dk) {
  constexpr int H = Halo<OP>::value;
  static_assert(OP == 2 || OP == 4, "diffusion only");
  const int ib = i0 & ~1;  // pairs start at even columns
  const int c0 = ib + 2 * (blockIdx.x * (32 * M2_WARPS) + threadIdx.x);
  if (c0 >= i0 + di) return;  // no shuffles and no barriers below
  const bool m0 = c0 >= i0, m1 = c0 + 1 < i0 + di;
  const int js = j0 + blockIdx.y * LJ, je = min(js + LJ, j0 + dj);
  const long long s1 = phi.s1;
  const int pmax = (int)((s1 - 2) & ~1LL);         // last aligned pair inside a row
  const int cl = max(c0 - 2, 0), cr = min(c0 + 2, pmax);  // neighbour pairs (clamped: only
                                                          // masked points see clamped values)
  const int jlast = phi.n1 - 1;
  for (int k = k0 + blockIdx.z; k < k0 + dk; k += gridDim.z) {
    const double *pc = phi.p + (c0 + k * phi.s2);  // row 0 of the own pair
    const double *pg = gam.p + (c0 * gam.s0 + js * gam.s1 + k * gam.s2);
    double *po = out.p + (c0 + js * out.s1 + k * out.s2);
    double2 w[2 * H + 1];  // rows j-H .. j+H of the own pair
#pragma unroll
    for (int m = 0; m < 2 * H; ++m) w[m + 1] = ldg2(pc + (js + m - H) * s1);
    double2 nxt = ldg2(pc + min(js + H, jlast) * s1);
    for (int j = js; j < je; ++j) {
#pragma unroll
      for (int m = 0; m < 2 * H; ++m) w[m] = w[m + 1];
      w[2 * H] = nxt;
      nxt = ldg2(pc + min(j + 1 + H, jlast) * s1);
      if (j + H + M2_PF < je + H)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + (j + H + M2_PF) * s1));
      const double *pr = phi.p + (j * s1 + k * phi.s2);
      const double2 l = ldg2(pr + cl), r = ldg2(pr + cr);
      const double g0 = __ldg(pg), g1 = __ldg(pg + gam.s0);
      double r0, r1;
      if (OP == 2) {
        r0 = lap2(g0, w[H].x, l.y, w[H].y, w[H - 1].x, w[H + 1].x, dx, dy);
        r1 = lap2(g1, w[H].y, w[H].x, r.x, w[H - 1].y, w[H + 1].y, dx, dy);
      } else {
        r0 = lap4(g0, w[H].x, l.x, l.y, w[H].y, r.x, w[0].x, w[H - 1].x, w[H + 1].x, w[2 * H].x, dx, dy);
        r1 = lap4(g1, w[H].y, l.y, w[H].x, r.x, r.y, w[0].y, w[H - 1].y, w[H + 1].y, w[2 * H].y, dx, dy);
      }
      if (m0 && m1) {
        if (!overwrite) {  // generics.py:L38-L40
          const double2 o = *reinterpret_cast<const double2 *>(po);
          r0 = o.x + r0;
          r1 = o.y + r1;
        }
        *reinterpret_cast<double2 *>(po) = make_double2(r0, r1);
      } else if (m0) {
        if (!overwrite) r0 = po[0] + r0;
        po[0] = r0;
      } else if (m1) {
        if (!overwrite) r1 = po[1] + r1;
        po[1] = r1;
      }
      pg += gam.s1;
      po += out.s1;
    }
  }
}

// ---- smoothing, two columns per lane: the marching structure of march2_kernel for the smoothers
// (OP 11 / 12 / 13 = first / second / third order, halo 1 / 2 / 3).  The tiled kernel the smoothers
// used to run pays two block barriers per level and reached 0.27 of the HBM peak on fields larger
// than L2 (profiles/README.md, round 2); the diffusion dwarf's marching kernel reaches 0.905.  Same
// point formulas in the same order as cross_kernel, hence the same bits.  The rim copy of
// HorizontalSmoothing.__call__ (rim_copy) is a separate frame launch.
// xm[d-1] / xp[d-1] = phi at i -/+ d, ym / yp likewise along j
template <int OP>
__device__ __forceinline__ double smooth_point(double g, double c, const double *xm, const double *xp,
                                               const double *ym, const double *yp) {
  if (OP == 11) {  // first_order.py:L124-L126
    return (1.0 - g) * c + 0.25 * g * (xm[0] + xp[0] + ym[0] + yp[0]);
  } else if (OP == 12) {  // second_order.py:L126-L139
    return (1.0 - 0.75 * g) * c +
           0.0625 * g * (-xm[1] + 4.0 * xm[0] - xp[1] + 4.0 * xp[0] - ym[1] + 4.0 * ym[0] - yp[1] + 4.0 * yp[0]);
  } else {  // third_order.py:L133-L150
    return (1.0 - 0.625 * g) * c +
           0.015625 * g *
               (xm[2] - 6.0 * xm[1] + 15.0 * xm[0] + xp[2] - 6.0 * xp[1] + 15.0 * xp[0] + ym[2] -
                6.0 * ym[1] + 15.0 * ym[0] + yp[2] - 6.0 * yp[1] + 15.0 * yp[0]);
  }
}

template <int OP, int LJ>
__global__ void __launch_bounds__(32 * M2_WARPS, 4)
    smooth2_kernel(View phi, View gam, View out, int i0, int j0, int k0, int di, int dj, int dk) {
  constexpr int H = Halo<OP>::value;
  static_assert(OP == 11 || OP == 12 || OP == 13, "smoothing only");
  const int ib = i0 & ~1;  // pairs start at even columns
  const int c0 = ib + 2 * (blockIdx.x * (32 * M2_WARPS) + threadIdx.x);
  if (c0 >= i0 + di) return;  // no shuffles and no barriers below
  const bool m0 = c0 >= i0, m1 = c0 + 1 < i0 + di;
  const int js = j0 + blockIdx.y * LJ, je = min(js + LJ, j0 + dj);
  const long long s1 = phi.s1;
  const int pmax = (int)((s1 - 2) & ~1LL);  // last aligned pair inside a row
  // neighbour pairs (clamped into the row: only masked points see clamped values)
  const int cl = max(c0 - 2, 0), cr = min(c0 + 2, pmax);
  const int cll = max(c0 - 4, 0), crr = min(c0 + 4, pmax);
  const int jlast = phi.n1 - 1;
  for (int k = k0 + blockIdx.z; k < k0 + dk; k += gridDim.z) {
    const double *pc = phi.p + (c0 + k * phi.s2);  // row 0 of the own pair
    const double *pg = gam.p + (c0 * gam.s0 + js * gam.s1 + k * gam.s2);
    double *po = out.p + (c0 + js * out.s1 + k * out.s2);
    double2 w[2 * H + 1];  // rows j-H .. j+H of the own pair
#pragma unroll
    for (int m = 0; m < 2 * H; ++m) w[m + 1] = ldg2(pc + max(js + m - H, 0) * s1);
    double2 nxt = ldg2(pc + min(js + H, jlast) * s1);
    for (int j = js; j < je; ++j) {
#pragma unroll
      for (int m = 0; m < 2 * H; ++m) w[m] = w[m + 1];
      w[2 * H] = nxt;
      nxt = ldg2(pc + min(j + 1 + H, jlast) * s1);
      if (j + H + M2_PF < je + H)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + (j + H + M2_PF) * s1));
      const double *pr = phi.p + (j * s1 + k * phi.s2);
      const double2 l = ldg2(pr + cl), r = ldg2(pr + cr);
      double2 ll = make_double2(0.0, 0.0), rr = ll;
      if (H == 3) {
        ll = ldg2(pr + cll);
        rr = ldg2(pr + crr);
      }
      const double g0 = __ldg(pg), g1 = __ldg(pg + gam.s0);
      // column c0: x offsets -1, -2, -3 = l.y, l.x, ll.y; +1, +2, +3 = own.y, r.x, r.y
      // column c1: x offsets -1, -2, -3 = own.x, l.y, l.x; +1, +2, +3 = r.x, r.y, rr.x
      const double xm0[3] = {l.y, l.x, ll.y}, xp0[3] = {w[H].y, r.x, r.y};
      const double xm1[3] = {w[H].x, l.y, l.x}, xp1[3] = {r.x, r.y, rr.x};
      double ym0[3], yp0[3], ym1[3], yp1[3];
#pragma unroll
      for (int d = 1; d <= 3; ++d) {
        if (d <= H) {
          ym0[d - 1] = w[H - d].x; yp0[d - 1] = w[H + d].x;
          ym1[d - 1] = w[H - d].y; yp1[d - 1] = w[H + d].y;
        } else {
          ym0[d - 1] = yp0[d - 1] = ym1[d - 1] = yp1[d - 1] = 0.0;
        }
      }
      const double r0 = smooth_point<OP>(g0, w[H].x, xm0, xp0, ym0, yp0);
      const double r1 = smooth_point<OP>(g1, w[H].y, xm1, xp1, ym1, yp1);
      if (m0 && m1) {
        *reinterpret_cast<double2 *>(po) = make_double2(r0, r1);
      } else if (m0) {
        po[0] = r0;
      } else if (m1) {
        po[1] = r1;
      }
      pg += gam.s1;
      po += out.s1;
    }
  }
}

template <int OP>
int launch_smooth2(const char *what, View phi, View gam, View out, const int32_t o[3], const int32_t d[3],
                   cudaStream_t st) {
  const int gz = d[2] > 65535 ? 65535 : d[2];
  const int cols = 2 * 32 * M2_WARPS;
  const int span = o[0] + d[0] - (o[0] & ~1);
  const int gx = (span + cols - 1) / cols;
  const long long blocks64 = (long long)gx * ((d[1] + 63) / 64) * gz;
  if (blocks64 >= 148 * 4) {
    dim3 grid(gx, (d[1] + 63) / 64, gz);
    smooth2_kernel<OP, 64><<<grid, 32 * M2_WARPS, 0, st>>>(phi, gam, out, o[0], o[1], o[2], d[0], d[1], d[2]);
  } else {
    dim3 grid(gx, (d[1] + 7) / 8, gz);
    smooth2_kernel<OP, 8><<<grid, 32 * M2_WARPS, 0, st>>>(phi, gam, out, o[0], o[1], o[2], d[0], d[1], d[2]);
  }
  return check_launch(what);
}

// the frame of the smoother's box around [o, o + d): out = phi (the four `copy` launches of
// HorizontalSmoothing.__call__).  One thread per FRAME point: full rows below and above the box,
// then the left and right margins of the rows in between.
__global__ void __launch_bounds__(256) rim_copy_kernel(View phi, View out, int i0, int j0, int k0, int di,
                                                       int dj, int dk, int ri, int rj) {
  const long long low = (long long)j0 * ri, high = (long long)(rj - j0 - dj) * ri;
  const int margin = ri - di;  // points per middle row outside the box
  const long long total = low + high + (long long)dj * margin;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  int i, j;
  if (t < low) {
    j = (int)(t / ri); i = (int)(t - (long long)j * ri);
  } else if (t < low + high) {
    const long long u = t - low;
    j = (int)(u / ri); i = (int)(u - (long long)j * ri);
    j += j0 + dj;
  } else {
    const long long u = t - low - high;
    j = (int)(u / margin);
    const int m = (int)(u - (long long)j * margin);
    i = m < i0 ? m : m + di;
    j += j0;
  }
  for (int k = k0 + blockIdx.y; k < k0 + dk; k += gridDim.y) out(i, j, k) = phi.ld(i, j, k);
}

int launch_rim_copy(View phi, View out, const int32_t o[3], const int32_t d[3], cudaStream_t st) {
  const int ri = 2 * o[0] + d[0], rj = 2 * o[1] + d[1];
  const long long total = (long long)ri * rj - (long long)d[0] * d[1];
  if (total <= 0 || d[2] <= 0) return TB200_OK;
  dim3 grid((unsigned)((total + 255) / 256), (unsigned)(d[2] > 65535 ? 65535 : d[2]), 1);
  rim_copy_kernel<<<grid, 256, 0, st>>>(phi, out, o[0], o[1], o[2], d[0], d[1], d[2], ri, rj);
  return check_launch("smoothing_rim");
}

// TB200_SMOOTH_IMPL: "march2" (default) or "tile"
int smooth_impl() {
  static int impl = -1;
  if (impl < 0) {
    const char *e = getenv("TB200_SMOOTH_IMPL");
    impl = (e != nullptr && strcmp(e, "tile") == 0) ? 0 : 2;
  }
  return impl;
}

bool march2_ok(const View &phi, const View &out) {
  const View *vs[] = {&phi, &out};
  for (const View *v : vs)
    if (v->s0 != 1 || (v->s1 & 1) != 0 || (v->s2 & 1) != 0 || (reinterpret_cast<uintptr_t>(v->p) & 15) != 0 ||
        v->s1 < 4)
      return false;
  return true;
}

// TB200_DIFF_IMPL: "march2" (default: two columns per lane), "march" (one column per lane,
// register windows along j) or "tile" (shared-memory tiles)
int diff_impl() {
  static int impl = -1;
  if (impl < 0) {
    const char *e = getenv("TB200_DIFF_IMPL");
    impl = e == nullptr ? 2 : strcmp(e, "march") == 0 ? 1 : strcmp(e, "tile") == 0 ? 0 : 2;
  }
  return impl;
}

template <int OP>
int launch_march2(const char *what, View phi, View gam, View out, const CDiv &cdx, const CDiv &cdy,
                  int overwrite, const int32_t o[3], const int32_t d[3], cudaStream_t st) {
  const int gz = d[2] > 65535 ? 65535 : d[2];
  const int cols = 2 * 32 * M2_WARPS;                       // columns per block
  const int span = o[0] + d[0] - (o[0] & ~1);               // columns from the first even one
  const int gx = (span + cols - 1) / cols;
  const long long blocks64 = (long long)gx * ((d[1] + 63) / 64) * gz;
  if (blocks64 >= 148 * 4) {  // enough 64-row strips to fill the machine
    dim3 grid(gx, (d[1] + 63) / 64, gz);
    march2_kernel<OP, 64><<<grid, 32 * M2_WARPS, 0, st>>>(phi, gam, out, cdx, cdy, overwrite, o[0], o[1],
                                                          o[2], d[0], d[1], d[2]);
  } else {
    dim3 grid(gx, (d[1] + 7) / 8, gz);
    march2_kernel<OP, 8><<<grid, 32 * M2_WARPS, 0, st>>>(phi, gam, out, cdx, cdy, overwrite, o[0], o[1],
                                                         o[2], d[0], d[1], d[2]);
  }
  return check_launch(what);
}

template <int OP>
int launch_march(const char *what, View phi, View gam, View out, const CDiv &cdx, const CDiv &cdy,
                 int overwrite, const int32_t o[3], const int32_t d[3], cudaStream_t st) {
  const int gz = d[2] > 65535 ? 65535 : d[2];
  const long long blocks64 = (long long)((d[0] + 127) / 128) * ((d[1] + 63) / 64) * gz;
  if (blocks64 >= 148 * 8) {  // enough 64-row strips to fill the machine
    dim3 grid((d[0] + 127) / 128, (d[1] + 63) / 64, gz);
    march_kernel<OP, 64><<<grid, 128, 0, st>>>(phi, gam, out, cdx, cdy, overwrite, o[0], o[1], o[2],
                                                d[0], d[1], d[2]);
  } else {
    dim3 grid((d[0] + 127) / 128, (d[1] + 15) / 16, gz);
    march_kernel<OP, 16><<<grid, 128, 0, st>>>(phi, gam, out, cdx, cdy, overwrite, o[0], o[1], o[2],
                                                d[0], d[1], d[2]);
  }
  return check_launch(what);
}

template <int OP>
int launch_cross(const char *what, View phi, View gam, View out, double dx, double dy,
                 int overwrite, int rim, const int32_t o[3], const int32_t d[3],
                 cudaStream_t st) {
  // rim box = the smoother's shape: nb + (n - 2 nb) + nb points a side
  const int ri = 2 * o[0] + d[0], rj = 2 * o[1] + d[1];
  const int ei = rim ? ri : d[0], ej = rim ? rj : d[1];
  if (ei <= 0 || ej <= 0 || d[2] <= 0) return TB200_OK;
  dim3 block(TX, TY, 1);
  dim3 grid((ei + TX - 1) / TX, (ej + TY - 1) / TY, d[2] > 65535 ? 65535 : d[2]);
  // full denominators, evaluated as the reference does: dx * dx and 12.0 * dx * dx
  const CDiv cdx = make_cdiv(OP == 4 ? 12.0 * dx * dx : (OP == 2 ? dx * dx : 1.0));
  const CDiv cdy = make_cdiv(OP == 4 ? 12.0 * dy * dy : (OP == 2 ? dy * dy : 1.0));
  if constexpr (OP == 2 || OP == 4) {
    if (!rim && diff_impl() == 2 && march2_ok(phi, out))
      return launch_march2<OP>(what, phi, gam, out, cdx, cdy, overwrite, o, d, st);
    if (!rim && diff_impl() >= 1) return launch_march<OP>(what, phi, gam, out, cdx, cdy, overwrite, o, d, st);
  }
  if constexpr (OP == 11 || OP == 12 || OP == 13) {
    // the marching kernel reads gamma with unit i-stride or as a broadcast; phi / out as pairs
    if (smooth_impl() == 2 && march2_ok(phi, out) && d[0] > 0 && d[1] > 0 && d[2] > 0 &&
        o[1] - Halo<OP>::value >= 0) {
      int rc = launch_smooth2<OP>(what, phi, gam, out, o, d, st);
      if (rc == TB200_OK && rim) rc = launch_rim_copy(phi, out, o, d, st);
      return rc;
    }
  }
  cross_kernel<OP><<<grid, block, 0, st>>>(phi, gam, out, cdx, cdy, overwrite, rim, o[0], o[1],
                                           o[2], d[0], d[1], d[2], ri, rj);
  return check_launch(what);
}


// ---- one-dimensional variants (second_order_1dx / _1dy, fourth_order_1dx / _1dy of the
// diffusers, first..third_order_1dx / _1dy of the smoothers): the stencil runs along ONE
// horizontal axis, for grids that are a single row or column of points.  Such grids are small
// (n x 1 x nz), so there is nothing to stage: one thread per point, neighbours straight from
// L1/L2.  The association of the reference's 1-D formulas differs from the 2-D ones
// (`gamma * (...) / (dx * dx)`: the product first), hence separate code rather than a flag.
template <int OP>
struct LineOp {
  View phi, gam, out;
  CDiv den;
  int si, sj;  // unit step along the stencil axis
  int overwrite, rim;
  int bi, bj, i0, j0, k0, di, dj;
  __device__ __forceinline__ double at(int i, int j, int k, int m) const {
    return phi.ld(i + m * si, j + m * sj, k);
  }
  __device__ void operator()(int ri, int rj, int rk) const {
    const int i = bi + ri, j = bj + rj, k = k0 + rk;
    const double c = phi.ld(i, j, k);
    const bool inside = i >= i0 && i < i0 + di && j >= j0 && j < j0 + dj;
    if (!inside) {
      if (rim) out(i, j, k) = c;  // the two `copy` launches of the 1-D smoothers' __call__
      return;
    }
    const double g = gam.ld(i, j, k);
    double r;
    if (OP == 2) {  // diffusers/second_order.py:L218, L329
      r = g * (at(i, j, k, -1) - 2.0 * c + at(i, j, k, 1)) / den;
    } else if (OP == 4) {  // diffusers/fourth_order.py:L265-L275, L402-L412
      r = g * (-at(i, j, k, -2) + 16.0 * at(i, j, k, -1) - 30.0 * c + 16.0 * at(i, j, k, 1) -
               at(i, j, k, 2)) / den;
    } else if (OP == 11) {  // smoothers/first_order.py:L216-L218, L308-L310
      r = (1.0 - 0.5 * g) * c + 0.25 * g * (at(i, j, k, -1) + at(i, j, k, 1));
    } else if (OP == 12) {  // smoothers/second_order.py:L240-L247, L344-L351
      r = (1.0 - 0.375 * g) * c +
          0.0625 * g * (-at(i, j, k, -2) + 4.0 * at(i, j, k, -1) - at(i, j, k, 2) + 4.0 * at(i, j, k, 1));
    } else {  // smoothers/third_order.py:L254-L263, L364-L373
      r = (1.0 - 0.3125 * g) * c +
          0.015625 * g * (at(i, j, k, -3) - 6.0 * at(i, j, k, -2) + 15.0 * at(i, j, k, -1) +
                          at(i, j, k, 3) - 6.0 * at(i, j, k, 2) + 15.0 * at(i, j, k, 1));
    }
    if ((OP == 2 || OP == 4) && !overwrite) r = out(i, j, k) + r;  // generics.py:L38-L40
    out(i, j, k) = r;
  }
};

template <int OP>
int launch_line(const char *what, View phi, View gam, View out, int axis, double h,
                int overwrite, int rim, const int32_t o[3], const int32_t d[3], cudaStream_t st) {
  // rim box = the smoother's shape along the stencil axis: nb + (n - 2 nb) + nb points
  const int ri = 2 * o[0] + d[0], rj = 2 * o[1] + d[1];
  LineOp<OP> op;
  op.phi = phi; op.gam = gam; op.out = out;
  // full denominators, evaluated as the reference does: h * h and 12.0 * h * h
  op.den = make_cdiv(OP == 4 ? 12.0 * h * h : (OP == 2 ? h * h : 1.0));
  op.si = axis == 0; op.sj = axis == 1;
  op.overwrite = overwrite; op.rim = rim;
  op.bi = rim ? 0 : o[0]; op.bj = rim ? 0 : o[1];
  op.i0 = o[0]; op.j0 = o[1]; op.k0 = o[2]; op.di = d[0]; op.dj = d[1];
  const int32_t ext[3] = {rim ? ri : d[0], rim ? rj : d[1], d[2]};
  return launch_box(what, ext, st, op);
}

// argument checks shared by the two 1-D entry points: halo h along `axis` only
int check_line(const char *what, const View &phi, const View &gam, const View &out, int axis,
               int h, const int32_t origin[3], const int32_t domain[3]) {
  TB200_REQUIRE(axis == 0 || axis == 1, "%s: axis must be 0 (x) or 1 (y) (got %d)", what, axis);
  const int hx = axis == 0 ? h : 0, hy = axis == 1 ? h : 0;
  TB200_REQUIRE(box_inside(phi, origin, domain, hx, hx, hy, hy),
                "%s: in_phi box + halo %d along axis %d outside storage", what, h, axis);
  TB200_REQUIRE(box_inside(gam, origin, domain) && box_inside(out, origin, domain),
                "%s: gamma/out box outside storage", what);
  TB200_REQUIRE(phi.p != out.p, "%s: in_phi and out_phi must not alias", what);
  return TB200_OK;
}

}  // namespace

extern "C" int tb200_diffusion(int order, const tb200_field *in_phi,
                               const tb200_field *in_gamma, tb200_field *out_phi, double dx,
                               double dy, int ow_out_phi, const int32_t origin[3],
                               const int32_t domain[3], void *stream) {
  View phi = view(in_phi), gam = view(in_gamma), out = view(out_phi);
  TB200_REQUIRE(order == 2 || order == 4, "diffusion: order must be 2 or 4 (got %d)", order);
  const int h = order / 2;
  TB200_REQUIRE(box_inside(phi, origin, domain, h, h, h, h),
                "diffusion: in_phi box + halo %d outside storage", h);
  TB200_REQUIRE(box_inside(gam, origin, domain) && box_inside(out, origin, domain),
                "diffusion: gamma/out box outside storage");
  TB200_REQUIRE(phi.p != out.p, "diffusion: in_phi and out_phi must not alias");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (order == 2)
    return launch_cross<2>("diffusion2", phi, gam, out, dx, dy, ow_out_phi, 0, origin, domain, st);
  return launch_cross<4>("diffusion4", phi, gam, out, dx, dy, ow_out_phi, 0, origin, domain, st);
}

extern "C" int tb200_smoothing(int order, const tb200_field *in_phi,
                               const tb200_field *in_gamma, tb200_field *out_phi,
                               int rim_copy, const int32_t origin[3], const int32_t domain[3],
                               void *stream) {
  View phi = view(in_phi), gam = view(in_gamma), out = view(out_phi);
  TB200_REQUIRE(order >= 1 && order <= 3, "smoothing: order must be 1..3 (got %d)", order);
  const int h = order;
  TB200_REQUIRE(box_inside(phi, origin, domain, h, h, h, h),
                "smoothing: in_phi box + halo %d outside storage", h);
  TB200_REQUIRE(box_inside(gam, origin, domain) && box_inside(out, origin, domain),
                "smoothing: gamma/out box outside storage");
  TB200_REQUIRE(phi.p != out.p, "smoothing: in_phi and out_phi must not alias");
  TB200_REQUIRE(!rim_copy || (2 * origin[0] + domain[0] <= phi.n0 && 2 * origin[0] + domain[0] <= out.n0 &&
                              2 * origin[1] + domain[1] <= phi.n1 && 2 * origin[1] + domain[1] <= out.n1),
                "smoothing: rim box outside storage");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (order == 1)
    return launch_cross<11>("smoothing1", phi, gam, out, 0, 0, 1, rim_copy, origin, domain, st);
  if (order == 2)
    return launch_cross<12>("smoothing2", phi, gam, out, 0, 0, 1, rim_copy, origin, domain, st);
  return launch_cross<13>("smoothing3", phi, gam, out, 0, 0, 1, rim_copy, origin, domain, st);
}

extern "C" int tb200_diffusion_1d(int order, int axis, const tb200_field *in_phi,
                                  const tb200_field *in_gamma, tb200_field *out_phi, double h,
                                  int ow_out_phi, const int32_t origin[3],
                                  const int32_t domain[3], void *stream) {
  View phi = view(in_phi), gam = view(in_gamma), out = view(out_phi);
  TB200_REQUIRE(order == 2 || order == 4, "diffusion_1d: order must be 2 or 4 (got %d)", order);
  if (int rc = check_line("diffusion_1d", phi, gam, out, axis, order / 2, origin, domain)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (order == 2)
    return launch_line<2>("diffusion2_1d", phi, gam, out, axis, h, ow_out_phi, 0, origin, domain, st);
  return launch_line<4>("diffusion4_1d", phi, gam, out, axis, h, ow_out_phi, 0, origin, domain, st);
}

extern "C" int tb200_smoothing_1d(int order, int axis, const tb200_field *in_phi,
                                  const tb200_field *in_gamma, tb200_field *out_phi,
                                  int rim_copy, const int32_t origin[3],
                                  const int32_t domain[3], void *stream) {
  View phi = view(in_phi), gam = view(in_gamma), out = view(out_phi);
  TB200_REQUIRE(order >= 1 && order <= 3, "smoothing_1d: order must be 1..3 (got %d)", order);
  if (int rc = check_line("smoothing_1d", phi, gam, out, axis, order, origin, domain)) return rc;
  TB200_REQUIRE(!rim_copy || (2 * origin[0] + domain[0] <= phi.n0 && 2 * origin[0] + domain[0] <= out.n0 &&
                              2 * origin[1] + domain[1] <= phi.n1 && 2 * origin[1] + domain[1] <= out.n1),
                "smoothing_1d: rim box outside storage");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (order == 1)
    return launch_line<11>("smoothing1_1d", phi, gam, out, axis, 0, 1, rim_copy, origin, domain, st);
  if (order == 2)
    return launch_line<12>("smoothing2_1d", phi, gam, out, axis, 0, 1, rim_copy, origin, domain, st);
  return launch_line<13>("smoothing3_1d", phi, gam, out, axis, 0, 1, rim_copy, origin, domain, st);
}

extern "C" int tb200_hyperdiffusion(const tb200_field *in_phi, tb200_field *out_phi, double alpha,
                                    const int32_t origin[3], const int32_t domain[3],
                                    void *stream) {
  View phi = view(in_phi), out = view(out_phi);
  TB200_REQUIRE(box_inside(phi, origin, domain, 3, 3, 3, 3),
                "hyperdiffusion: in_phi box + halo 3 outside storage");
  TB200_REQUIRE(box_inside(out, origin, domain), "hyperdiffusion: out box outside storage");
  TB200_REQUIRE(phi.p != out.p, "hyperdiffusion: in_phi and out_phi must not alias");
  HyperOp op;
  op.phi = phi; op.out = out; op.alpha = alpha;
  op.i0 = origin[0]; op.j0 = origin[1]; op.k0 = origin[2];
  return launch_box("hyperdiffusion", domain, static_cast<cudaStream_t>(stream), op);
}
