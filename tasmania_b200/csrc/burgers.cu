// burgers.cu -- K10: the forward-Euler stage of the 2-D Burgers dwarf with u and v fused
// in one kernel (src/tasmania/burgers/dynamics/stepper.py:L188-L227; advection subroutines
// burgers/dynamics/subclasses/advection/{first..sixth}_order.py).
//
// Roofline: HBM, 48 B/point/stage (read u, v now + u, v provisional, write u, v).  The
// reference grid has nz == 1 (burgers/dynamics/dycore.py:L100), i.e. 10^4 points: launch
// latency bound; the kernel is a plain one-thread-per-point cross stencil whose halo
// re-reads are served by L1/L2.
#include "stencil_math.cuh"

using namespace tb200;

namespace {
template <int ORDER>
int run(View u, View v, View ut, View vt, View ou, View ov, View tu, View tv, double dt,
        double dx, double dy, const int32_t o[3], const int32_t d[3], cudaStream_t st) {
  using A = Advection<ORDER>;
  const int i0 = o[0], j0 = o[1], k0 = o[2];
  const CDiv denx = make_cdiv(A::denominator(dx)), deny = make_cdiv(A::denominator(dy));
  return launch_box("burgers_forward_euler", d, st, [=] __device__(int i, int j, int k) {
    i += i0; j += j0; k += k0;
    const double *pu = ut.p + (i * ut.s0 + j * ut.s1 + k * ut.s2);
    const double *pv = vt.p + (i * vt.s0 + j * vt.s1 + k * vt.s2);
    const double a = __ldg(pu) / denx, b = __ldg(pv) / deny;
    const double adv_u_x = A::term(a, pu, ut.s0);
    const double adv_u_y = A::term(b, pu, ut.s1);
    const double adv_v_x = A::term(a, pv, vt.s0);
    const double adv_v_y = A::term(b, pv, vt.s1);
    // stepper.py:L219-L227
    if (tu.ok())
      ou(i, j, k) = u(i, j, k) - dt * (adv_u_x + adv_u_y - tu(i, j, k));
    else
      ou(i, j, k) = u(i, j, k) - dt * (adv_u_x + adv_u_y);
    if (tv.ok())
      ov(i, j, k) = v(i, j, k) - dt * (adv_v_x + adv_v_y - tv(i, j, k));
    else
      ov(i, j, k) = v(i, j, k) - dt * (adv_v_x + adv_v_y);
  });
}
}  // namespace

extern "C" int tb200_burgers_forward_euler(int advection_order, const tb200_field *in_u,
                                           const tb200_field *in_v, const tb200_field *in_u_tmp,
                                           const tb200_field *in_v_tmp, tb200_field *out_u,
                                           tb200_field *out_v, const tb200_field *in_u_tnd,
                                           const tb200_field *in_v_tnd, double dt, double dx,
                                           double dy, const int32_t origin[3],
                                           const int32_t domain[3], void *stream) {
  TB200_REQUIRE(advection_order >= 1 && advection_order <= 6,
                "burgers_forward_euler: advection order must be 1..6 (got %d)", advection_order);
  const int e = (advection_order + 1) / 2;
  View u = view(in_u), v = view(in_v), ut = view(in_u_tmp), vt = view(in_v_tmp);
  View ou = view(out_u), ov = view(out_v), tu = view(in_u_tnd), tv = view(in_v_tnd);
  TB200_REQUIRE(box_inside(u, origin, domain) && box_inside(v, origin, domain) &&
                    box_inside(ou, origin, domain) && box_inside(ov, origin, domain),
                "burgers_forward_euler: box outside storage");
  TB200_REQUIRE(box_inside(ut, origin, domain, e, e, e, e) && box_inside(vt, origin, domain, e, e, e, e),
                "burgers_forward_euler: provisional fields box + extent %d outside storage", e);
  TB200_REQUIRE((!tu.ok() || box_inside(tu, origin, domain)) && (!tv.ok() || box_inside(tv, origin, domain)),
                "burgers_forward_euler: tendency box outside storage");
  TB200_REQUIRE(ut.p != ou.p && vt.p != ov.p && ut.p != ov.p && vt.p != ou.p,
                "burgers_forward_euler: provisional and output fields must not alias");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (advection_order) {
    case 1: return run<1>(u, v, ut, vt, ou, ov, tu, tv, dt, dx, dy, origin, domain, st);
    case 2: return run<2>(u, v, ut, vt, ou, ov, tu, tv, dt, dx, dy, origin, domain, st);
    case 3: return run<3>(u, v, ut, vt, ou, ov, tu, tv, dt, dx, dy, origin, domain, st);
    case 4: return run<4>(u, v, ut, vt, ou, ov, tu, tv, dt, dx, dy, origin, domain, st);
    case 5: return run<5>(u, v, ut, vt, ou, ov, tu, tv, dt, dx, dy, origin, domain, st);
    default: return run<6>(u, v, ut, vt, ou, ov, tu, tv, dt, dx, dy, origin, domain, st);
  }
}
