// stage.cuh -- argument block and addressing helpers shared by the fused-stage kernels
// (isentropic_fused.cu: thread-per-column / register-window kernels; isentropic_tma.cu: the
// TMA + shared-memory row-pipeline kernels).
#pragma once
#include "stencil_math.cuh"

namespace tb200 {

struct StageArgs {
  View s_now, su_now, sv_now, mtg_now;
  View s_int, su_int, sv_int, u_int, v_int;
  View s_new, su_new, sv_new, u_new, v_new;
  View s_ref, su_ref, sv_ref, u_ref, v_ref;
  View gamma, rmat, hs, exn, mtg, spre;
  View s_tnd, su_tnd, sv_tnd;  // slow tendencies (tb200_isentropic_stage.s_tnd ...), NULL = none
  // moist stage: mass fractions of the water constituents (qv, qc, qr), ntr = 0 when dry
  View q_now[3], q_int[3], q_new[3], q_ref[3];
  int ntr;
  int nx, ny, nz, nb, damp;
  int bx0, by0;    // block offsets of a partial launch of the momentum kernel
  int part, rim[4];  // tb200_isentropic_stage.part / .rim
  int derive_uv, skip_uv;  // tb200_isentropic_stage.derive_uv_in / .skip_uv_out
  int a2_ok;               // the layout allows the two-column kernels (16-byte aligned pairs)
  int periodic;            // tb200_isentropic_stage.periodic: wrap s_pre between the s-step and the scans
  double dt, dt_full, dx, dy, dz, eps, pt, theta_s, pref, rd, g, cp;
  FluxConst fc;
  CDiv two_dx, two_dy, cpref;
};

// All 3-D fields of a fused stage share one geometry (unit i-stride, equal row and plane
// strides -- what the b200 allocator produces for equal shapes; checked on the host), so one
// 32-bit running BYTE offset addresses every field: loads compile to
// [uniform base + offset + immediate] with no per-load integer arithmetic.
__device__ __forceinline__ double ldo(const double *base, unsigned off) {
  return __ldg(reinterpret_cast<const double *>(reinterpret_cast<const char *>(base) + off));
}
__device__ __forceinline__ void sto(double *base, unsigned off, double v) {
  *reinterpret_cast<double *>(reinterpret_cast<char *>(base) + off) = v;
}
__device__ __forceinline__ const double *ptr_at(const double *base, unsigned off) {
  return reinterpret_cast<const double *>(reinterpret_cast<const char *>(base) + off);
}
// pull a line towards L2 ahead of its use (no register, no stall)
__device__ __forceinline__ void prefetch_l2(const double *base, unsigned off) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(base) + off));
}


// implemented in isentropic_tma.cu: the TMA row-pipeline kernels of the stage.  Each returns
// TB200_OK, an error code, or -1 when this configuration is not covered (the caller then runs
// the register-window kernel of isentropic_fused.cu).
int launch_stage_c(const StageArgs &a, int scheme, cudaStream_t st);

}  // namespace tb200
