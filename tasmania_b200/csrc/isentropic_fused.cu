// isentropic_fused.cu -- one Runge-Kutta stage of the dry isentropic dynamical core with the
// relaxed lateral boundary, fused into three kernels (the benchmark hot path).
//
// Reference sequence per stage (src/tasmania/isentropic/dynamics/dycore.py:L641-L721 and
// subclasses/prognostics/rk3ws_si.py:L105-L234):
//   K1 step s -> irelax(s) -> montgomery(s_new) -> K2 step su, sv -> irelax(s, su, sv, u, v)
//   -> Rayleigh damping (s, su, sv) -> velocity_x / velocity_y -> outermost layers of u, v.
// The reference runs 13 full-domain passes for this; here:
//
//   kernel S  (one thread per column, marching in k)
//       s_pre = irelax(K1(...))                      -> s_new
//       p, exn by the downward scan on s_pre         -> scratch_exn   (exn at interface k+1)
//       mtg_new by the upward scan                   -> scratch_mtg
//   kernel M  (one thread per point)
//       su, sv = irelax(K2(...)), s = irelax(s_pre), then Rayleigh damping on all three
//   kernel V  (one thread per point)
//       u, v from the final s, su, sv; outermost faces from the reference state.
//
// The vertical scans are inherently two sweeps (pressure top-down, Montgomery bottom-up) and
// the momentum step needs mtg_new at i+-1 / j+-1, hence the kernel boundary between S and M;
// V needs the *final* su/s at i-1 / j-1, hence the boundary between M and V.
//
// HBM traffic per point and stage (8-byte words): S reads s_now, s_int, u, v, writes s_new,
// exn, re-reads exn (L2-resident for the CTA's columns at moderate sizes), writes mtg
// = 7-8 words; M reads s_now, s_new, mtg_now, mtg_new, u, v, su_now, su_int, sv_now, sv_int,
// writes s, su, sv = 13 words; V reads s, su, sv, writes u, v = 5 words.  Total ~25-26 words
// = 200-208 B against the algorithmic minimum of 112 B; the relaxation
// band and the damping layer add reads of the reference fields only where gamma != 0 or
// R != 0.  All arithmetic follows the reference's operation order (see stencil_math.cuh).
#include "stencil_math.cuh"

using namespace tb200;

namespace {

struct StageArgs {
  View s_now, su_now, sv_now, mtg_now;
  View s_int, su_int, sv_int, u_int, v_int;
  View s_new, su_new, sv_new, u_new, v_new;
  View s_ref, su_ref, sv_ref, u_ref, v_ref;
  View gamma, rmat, hs, exn, mtg;
  int nx, ny, nz, nb, damp;
  double dt, dt_full, dx, dy, dz, eps, pt, theta_s, pref, rd, g, cp;
  FluxConst fc;
  CDiv two_dx, two_dy, cpref;
};

// ---------------------------------------------------------------- kernel S
template <int SCHEME>
__global__ void __launch_bounds__(128) stage_s_kernel(const StageArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= a.nx || j >= a.ny) return;
  const bool interior = i >= a.nb && i < a.nx - a.nb && j >= a.nb && j < a.ny - a.nb;
  const double gam = a.gamma.ld(i, j, 0);
  const double kappa = a.rd / a.cp;
  const double gdz = a.g * a.dz;

  // downward sweep: step s, relax, integrate the pressure, park exn[k+1] in scratch_exn[k]
  double p = a.pt;
  for (int k = 0; k < a.nz; ++k) {
    double v;
    if (interior) {
      const FaceVel w = face_velocities<SCHEME>(a.u_int, a.v_int, i, j, k, a.fc);
      const double div = flux_divergence<SCHEME>(w, a.s_int, i, j, k, a.fc);
      v = a.s_now.ld(i, j, k) - a.dt * (div - 0.0);
    } else {
      v = a.s_new(i, j, k);  // untouched by K1; the relaxation below decides
    }
    if (gam != 0.0) v = relax_point(gam, v, a.s_ref.ld(i, j, k));
    a.s_new(i, j, k) = v;
    p = p + gdz * v;
    a.exn(i, j, k) = a.cp * pow(p / a.cpref, kappa);
  }
  // upward sweep, diagnostics.py:L433-L438
  const double ex_s = a.exn(i, j, a.nz - 1);
  const double mtg_s = a.theta_s * ex_s + a.g * a.hs.ld(i, j, 0);
  double m = mtg_s + 0.5 * a.dz * ex_s;
  a.mtg(i, j, a.nz - 1) = m;
  for (int k = a.nz - 2; k >= 0; --k) {
    m = m + a.dz * a.exn(i, j, k);
    a.mtg(i, j, k) = m;
  }
}

// ---------------------------------------------------------------- kernel M
template <int SCHEME>
__global__ void __launch_bounds__(256) stage_m_kernel(const StageArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i >= a.nx || j >= a.ny) return;
  const bool interior = i >= a.nb && i < a.nx - a.nb && j >= a.nb && j < a.ny - a.nb;
  const double gam = a.gamma.ld(i, j, 0);
  const double r = a.damp ? a.rmat.ld(0, 0, k) : 0.0;

  double s = a.s_new(i, j, k);  // s_pre written by kernel S
  double su, sv;
  const bool need_now = interior || r != 0.0;
  const double s_now = need_now ? a.s_now.ld(i, j, k) : 0.0;
  const double su_now = need_now ? a.su_now.ld(i, j, k) : 0.0;
  const double sv_now = need_now ? a.sv_now.ld(i, j, k) : 0.0;
  if (interior) {
    const FaceVel w = face_velocities<SCHEME>(a.u_int, a.v_int, i, j, k, a.fc);
    // prognostics/utils.py:L191-L204
    {
      const double div = flux_divergence<SCHEME>(w, a.su_int, i, j, k, a.fc);
      const double pg_now = (1.0 - a.eps) * s_now *
                            (a.mtg_now.ld(i + 1, j, k) - a.mtg_now.ld(i - 1, j, k)) / a.two_dx;
      const double pg_new =
          a.eps * s * (a.mtg.ld(i + 1, j, k) - a.mtg.ld(i - 1, j, k)) / a.two_dx;
      su = su_now - a.dt * (div + pg_now + pg_new - 0.0);
    }
    {
      const double div = flux_divergence<SCHEME>(w, a.sv_int, i, j, k, a.fc);
      const double pg_now = (1.0 - a.eps) * s_now *
                            (a.mtg_now.ld(i, j + 1, k) - a.mtg_now.ld(i, j - 1, k)) / a.two_dy;
      const double pg_new =
          a.eps * s * (a.mtg.ld(i, j + 1, k) - a.mtg.ld(i, j - 1, k)) / a.two_dy;
      sv = sv_now - a.dt * (div + pg_now + pg_new - 0.0);
    }
  } else {
    su = a.su_new(i, j, k);
    sv = a.sv_new(i, j, k);
  }
  const bool need_ref = gam != 0.0 || r != 0.0;
  const double s_ref = need_ref ? a.s_ref.ld(i, j, k) : 0.0;
  const double su_ref = need_ref ? a.su_ref.ld(i, j, k) : 0.0;
  const double sv_ref = need_ref ? a.sv_ref.ld(i, j, k) : 0.0;
  if (gam != 0.0) {  // hb.enforce_raw, dycore.py:L686
    s = relax_point(gam, s, s_ref);
    su = relax_point(gam, su, su_ref);
    sv = relax_point(gam, sv, sv_ref);
  }
  if (r != 0.0) {  // dycore.py:L694-L700 (R == 0 leaves the value unchanged bit for bit)
    s = damp_point(s_now, s, s_ref, r, a.dt_full);
    su = damp_point(su_now, su, su_ref, r, a.dt_full);
    sv = damp_point(sv_now, sv, sv_ref, r, a.dt_full);
  }
  a.s_new(i, j, k) = s;
  a.su_new(i, j, k) = su;
  a.sv_new(i, j, k) = sv;
}

// ---------------------------------------------------------------- kernel V
__global__ void __launch_bounds__(256) stage_v_kernel(const StageArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z;
  if (i > a.nx || j > a.ny) return;
  // dwarfs/diagnostics.py:L219-L272 + relaxed.py:L161-L191
  const bool in_i = i < a.nx, in_j = j < a.ny;
  double s = 0.0, su = 0.0, sv = 0.0;
  if (in_i && in_j) {
    s = a.s_new(i, j, k);
    su = a.su_new(i, j, k);
    sv = a.sv_new(i, j, k);
  }
  if (in_j) {
    double u;
    if (i == 0 || i == a.nx) {
      u = a.u_ref.ld(i, j, k);
    } else {
      u = (a.su_new(i - 1, j, k) + su) / (a.s_new(i - 1, j, k) + s);
    }
    a.u_new(i, j, k) = u;
  }
  if (in_i) {
    double v;
    if (j == 0 || j == a.ny) {
      v = a.v_ref.ld(i, j, k);
    } else {
      v = (a.sv_new(i, j - 1, k) + sv) / (a.s_new(i, j - 1, k) + s);
    }
    a.v_new(i, j, k) = v;
  }
}

template <int SCHEME>
int run_stage(const StageArgs &a, cudaStream_t st) {
  {
    dim3 block(32, 4, 1);
    dim3 grid((a.nx + 31) / 32, (a.ny + 3) / 4, 1);
    stage_s_kernel<SCHEME><<<grid, block, 0, st>>>(a);
    int rc = check_launch("isentropic_stage_dry/S");
    if (rc) return rc;
  }
  {
    dim3 block(64, 4, 1);
    dim3 grid((a.nx + 63) / 64, (a.ny + 3) / 4, a.nz);
    stage_m_kernel<SCHEME><<<grid, block, 0, st>>>(a);
    int rc = check_launch("isentropic_stage_dry/M");
    if (rc) return rc;
  }
  {
    dim3 block(64, 4, 1);
    dim3 grid((a.nx + 1 + 63) / 64, (a.ny + 1 + 3) / 4, a.nz);
    stage_v_kernel<<<grid, block, 0, st>>>(a);
    return check_launch("isentropic_stage_dry/V");
  }
}

bool covers(const View &v, int ni, int nj, int nk) {
  return v.ok() && v.n0 >= ni && v.n1 >= nj && v.n2 >= nk;
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
    tb200_field *scratch_exn, tb200_field *scratch_mtg, void *stream) {
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
  a.exn = view(scratch_exn); a.mtg = view(scratch_mtg);
  a.nx = cfg->nx; a.ny = cfg->ny; a.nz = cfg->nz; a.nb = cfg->nb; a.damp = cfg->damp;
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
                        &a.sv_ref, &a.exn,    &a.mtg};
  for (const View *v : mass)
    TB200_REQUIRE(covers(*v, nx, ny, nz), "isentropic_stage_dry: a mass-point field is NULL or too small");
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
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (cfg->flux_scheme) {
    case TB200_FLUX_UPWIND: return run_stage<TB200_FLUX_UPWIND>(a, st);
    case TB200_FLUX_CENTERED: return run_stage<TB200_FLUX_CENTERED>(a, st);
    case TB200_FLUX_THIRD_ORDER_UPWIND: return run_stage<TB200_FLUX_THIRD_ORDER_UPWIND>(a, st);
    default: return run_stage<TB200_FLUX_FIFTH_ORDER_UPWIND>(a, st);
  }
}
