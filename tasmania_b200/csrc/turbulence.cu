// turbulence.cu -- Smagorinsky horizontal turbulence.
//
// Reference (numpy definitions):
//   src/tasmania/physics/turbulence.py:L165-L229            Smagorinsky2d (u, v -> tendencies)
//   src/tasmania/isentropic/physics/turbulence.py:L99-L125  IsentropicSmagorinsky (s, su, sv)
//
//   s00 = du/dx, s01 = (du/dy + dv/dx) / 2, s11 = dv/dy        (centred, over 2 dx / 2 dy)
//   nu  = cs^2 dx dy sqrt(2 (s00^2 + 2 s01^2 + s11^2))
//   u_tnd = 2 (d(nu s00)/dx + d(nu s01)/dy),  v_tnd = 2 (d(nu s01)/dx + d(nu s11)/dy)
// and, for the isentropic model, u = su / s, v = sv / s on the way in and s u_tnd, s v_tnd on
// the way out.  The footprint is a 13-point diamond of radius 2; a block stages it in shared memory
// in three steps -- velocities on the tile + 2, nu s00 / nu s01 / nu s11 on the tile + 1,
// tendencies on the tile -- so every division of su / s and every square root is evaluated once
// per point (plus the halo), and global memory is read once: 24 + 16 B/point (isentropic).
// Operation order is the reference's, divisions (by 2 dx / 2 dy through CDiv, by s through qdiv:
// both correctly rounded, common.cuh) and the square root are IEEE: bit-identical.
#include <math.h>

#include "common.cuh"

using namespace tb200;

namespace {

constexpr int TX = 32, TY = 8;             // outputs per block
constexpr int VX = TX + 4, VY = TY + 4;    // velocities: tile + 2
constexpr int PX = TX + 2, PY = TY + 2;    // products:   tile + 1

struct SmagArgs {
  View s, a, b;        // isentropic: s, su, sv; otherwise a = u, b = v
  View out_a, out_b;
  View base_a, base_b;  // stepped mode: out = base + factor * tendency (one stage of a tendency stepper)
  double factor;
  CDiv two_dx, two_dy;  // 2 dx, 2 dy (correctly rounded division by a constant: same bits as `/`)
  double coeff;         // cs^2 dx dy
  bool ow_a, ow_b;
  int i0, j0, k0, di, dj, dk;
};

template <bool ISEN, bool STEP>
__global__ void __launch_bounds__(TX *TY) smagorinsky_kernel(const SmagArgs a) {
  __shared__ double u[VY][VX + 1], v[VY][VX + 1];
  __shared__ double p00[PY][PX + 1], p01[PY][PX + 1], p11[PY][PX + 1];
  const int tid = threadIdx.y * TX + threadIdx.x;
  const int bi = a.i0 + blockIdx.x * TX, bj = a.j0 + blockIdx.y * TY;  // first output of the block
  const int ie = a.i0 + a.di, je = a.j0 + a.dj;                        // end of the box
  for (int k = a.k0 + blockIdx.z; k < a.k0 + a.dk; k += gridDim.z) {
    // ---- velocities on the tile + 2 (clipped to box + 2, which lies inside the storages)
    for (int n = tid; n < VX * VY; n += TX * TY) {
      const int lx = n % VX, ly = n / VX;
      const int i = bi - 2 + lx, j = bj - 2 + ly;
      double uu = 0.0, vv = 0.0;
      if (i < ie + 2 && j < je + 2) {
        if (ISEN) {  // isentropic/physics/turbulence.py:L119-L120
          const double sd = a.s.ld(i, j, k);
          uu = qdiv(a.a.ld(i, j, k), sd);  // same bits as `/` (zero momenta keep the fast path)
          vv = qdiv(a.b.ld(i, j, k), sd);
        } else {
          uu = a.a.ld(i, j, k);
          vv = a.b.ld(i, j, k);
        }
      }
      u[ly][lx] = uu;
      v[ly][lx] = vv;
    }
    __syncthreads();
    // ---- strain rates and nu on the tile + 1, turbulence.py:L212-L218
    for (int n = tid; n < PX * PY; n += TX * TY) {
      const int lx = n % PX, ly = n / PX;
      const int cx = lx + 1, cy = ly + 1;  // position in the velocity tile
      const double s00 = (u[cy][cx + 1] - u[cy][cx - 1]) / a.two_dx;
      const double s01 = 0.5 * ((u[cy + 1][cx] - u[cy - 1][cx]) / a.two_dy +
                                (v[cy][cx + 1] - v[cy][cx - 1]) / a.two_dx);
      const double s11 = (v[cy + 1][cx] - v[cy - 1][cx]) / a.two_dy;
      const double nu = a.coeff * sqrt(2.0 * (s00 * s00 + 2.0 * (s01 * s01) + s11 * s11));
      p00[ly][lx] = nu * s00;
      p01[ly][lx] = nu * s01;
      p11[ly][lx] = nu * s11;
    }
    __syncthreads();
    // ---- tendencies on the tile, turbulence.py:L219-L226
    {
      const int i = bi + threadIdx.x, j = bj + threadIdx.y;
      if (i < ie && j < je) {
        const int cx = threadIdx.x + 1, cy = threadIdx.y + 1;
        double ta = 2.0 * ((p00[cy][cx + 1] - p00[cy][cx - 1]) / a.two_dx +
                           (p01[cy + 1][cx] - p01[cy - 1][cx]) / a.two_dy);
        double tb = 2.0 * ((p01[cy][cx + 1] - p01[cy][cx - 1]) / a.two_dx +
                           (p11[cy + 1][cx] - p11[cy - 1][cx]) / a.two_dy);
        if (ISEN) {  // isentropic/physics/turbulence.py:L122-L123
          const double sd = a.s.ld(i, j, k);
          ta = sd * ta;
          tb = sd * tb;
        }
        double &oa = a.out_a(i, j, k), &ob = a.out_b(i, j, k);
        if (STEP) {  // the stage update of the tendency stepper (fma: a + f * b)
          oa = a.base_a.ld(i, j, k) + a.factor * ta;
          ob = a.base_b.ld(i, j, k) + a.factor * tb;
        } else {
          oa = a.ow_a ? ta : oa + ta;  // set_output
          ob = a.ow_b ? tb : ob + tb;
        }
      }
    }
    __syncthreads();  // the tiles are rewritten by the next level
  }
}

}  // namespace

extern "C" int tb200_smagorinsky(const tb200_field *in_s, const tb200_field *in_a,
                                 const tb200_field *in_b, tb200_field *out_a_tnd,
                                 tb200_field *out_b_tnd, double dx, double dy, double cs,
                                 int ow_out_a_tnd, int ow_out_b_tnd, const int32_t origin[3],
                                 const int32_t domain[3], void *stream) {
  SmagArgs a{};
  a.s = view(in_s); a.a = view(in_a); a.b = view(in_b);
  a.out_a = view(out_a_tnd); a.out_b = view(out_b_tnd);
  const bool isen = a.s.ok();
  a.two_dx = make_cdiv(2.0 * dx); a.two_dy = make_cdiv(2.0 * dy);
  a.coeff = pow(cs, 2.0) * dx * dy;  // cs**2 * dx * dy exactly as Python evaluates it (libm pow)
  a.ow_a = ow_out_a_tnd != 0; a.ow_b = ow_out_b_tnd != 0;
  a.i0 = origin[0]; a.j0 = origin[1]; a.k0 = origin[2];
  a.di = domain[0]; a.dj = domain[1]; a.dk = domain[2];
  TB200_REQUIRE(box_inside(a.a, origin, domain, 2, 2, 2, 2) && box_inside(a.b, origin, domain, 2, 2, 2, 2) &&
                    (!isen || box_inside(a.s, origin, domain, 2, 2, 2, 2)),
                "smagorinsky: box + 2 outside an input storage (needs nb >= 2)");
  TB200_REQUIRE(box_inside(a.out_a, origin, domain) && box_inside(a.out_b, origin, domain),
                "smagorinsky: box outside an output storage");
  TB200_REQUIRE(a.out_a.p != a.a.p && a.out_a.p != a.b.p && a.out_b.p != a.a.p && a.out_b.p != a.b.p &&
                    a.out_a.p != a.s.p && a.out_b.p != a.s.p && a.out_a.p != a.out_b.p,
                "smagorinsky: outputs must not alias the inputs or each other");
  if (a.di <= 0 || a.dj <= 0 || a.dk <= 0) return TB200_OK;
  dim3 block(TX, TY, 1);
  dim3 grid((a.di + TX - 1) / TX, (a.dj + TY - 1) / TY, a.dk > 65535 ? 65535 : a.dk);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (isen)
    smagorinsky_kernel<true, false><<<grid, block, 0, st>>>(a);
  else
    smagorinsky_kernel<false, false><<<grid, block, 0, st>>>(a);
  return check_launch("smagorinsky");
}

// Smagorinsky fused with the stage update of a tendency stepper (b200 only):
//   out = base + factor * tendency on [origin, origin + domain), out = base + factor * 0 on the rest
// of the `full` box (the reference path runs `fma` over the whole storages, whose tendency entries
// outside the box are zero).  The frame is one small launch of its own.
namespace {
__global__ void __launch_bounds__(256) frame_fma0_kernel(View ba, View bb, View oa, View ob, double factor,
                                                         int i0, int j0, int k0, int di, int dj, int dk,
                                                         int fi, int fj, int fk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= fi || j >= fj) return;
  const bool col_in = i >= i0 && i < i0 + di && j >= j0 && j < j0 + dj;
  const double z = factor * 0.0;
  for (int k = blockIdx.z; k < fk; k += gridDim.z) {
    if (col_in && k >= k0 && k < k0 + dk) continue;
    oa(i, j, k) = ba.ld(i, j, k) + z;
    ob(i, j, k) = bb.ld(i, j, k) + z;
  }
}
}  // namespace

extern "C" int tb200_smagorinsky_step(const tb200_field *in_s, const tb200_field *in_a,
                                      const tb200_field *in_b, const tb200_field *base_a,
                                      const tb200_field *base_b, tb200_field *out_a, tb200_field *out_b,
                                      double dx, double dy, double cs, double factor,
                                      const int32_t origin[3], const int32_t domain[3],
                                      const int32_t full[3], void *stream) {
  SmagArgs a{};
  a.s = view(in_s); a.a = view(in_a); a.b = view(in_b);
  a.out_a = view(out_a); a.out_b = view(out_b);
  a.base_a = view(base_a); a.base_b = view(base_b);
  a.factor = factor;
  const bool isen = a.s.ok();
  a.two_dx = make_cdiv(2.0 * dx); a.two_dy = make_cdiv(2.0 * dy);
  a.coeff = pow(cs, 2.0) * dx * dy;
  a.ow_a = a.ow_b = true;
  a.i0 = origin[0]; a.j0 = origin[1]; a.k0 = origin[2];
  a.di = domain[0]; a.dj = domain[1]; a.dk = domain[2];
  const int32_t zero[3] = {0, 0, 0};
  TB200_REQUIRE(box_inside(a.a, origin, domain, 2, 2, 2, 2) && box_inside(a.b, origin, domain, 2, 2, 2, 2) &&
                    (!isen || box_inside(a.s, origin, domain, 2, 2, 2, 2)),
                "smagorinsky_step: box + 2 outside an input storage (needs nb >= 2)");
  TB200_REQUIRE(box_inside(a.out_a, zero, full) && box_inside(a.out_b, zero, full) &&
                    box_inside(a.base_a, zero, full) && box_inside(a.base_b, zero, full),
                "smagorinsky_step: full box outside a base / output storage");
  TB200_REQUIRE(origin[0] >= 0 && origin[1] >= 0 && origin[2] >= 0 && origin[0] + domain[0] <= full[0] &&
                    origin[1] + domain[1] <= full[1] && origin[2] + domain[2] <= full[2],
                "smagorinsky_step: the tendency box must lie inside the full box");
  TB200_REQUIRE(a.out_a.p != a.a.p && a.out_a.p != a.b.p && a.out_b.p != a.a.p && a.out_b.p != a.b.p &&
                    a.out_a.p != a.s.p && a.out_b.p != a.s.p && a.out_a.p != a.out_b.p,
                "smagorinsky_step: outputs must not alias the inputs or each other");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (full[0] > 0 && full[1] > 0 && full[2] > 0) {
    dim3 block(64, 4, 1);
    dim3 grid((full[0] + 63) / 64, (full[1] + 3) / 4, full[2] > 64 ? 64 : full[2]);
    frame_fma0_kernel<<<grid, block, 0, st>>>(a.base_a, a.base_b, a.out_a, a.out_b, factor, a.i0, a.j0, a.k0,
                                              a.di, a.dj, a.dk, full[0], full[1], full[2]);
    if (int rc = check_launch("smagorinsky_step(frame)")) return rc;
  }
  if (a.di <= 0 || a.dj <= 0 || a.dk <= 0) return TB200_OK;
  dim3 block(TX, TY, 1);
  dim3 grid((a.di + TX - 1) / TX, (a.dj + TY - 1) / TY, a.dk > 65535 ? 65535 : a.dk);
  if (isen)
    smagorinsky_kernel<true, true><<<grid, block, 0, st>>>(a);
  else
    smagorinsky_kernel<false, true><<<grid, block, 0, st>>>(a);
  return check_launch("smagorinsky_step");
}
