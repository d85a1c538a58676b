// elementwise.cu -- point-wise stencils: K12 (copy/math/sts), K5 (relax), K6 (Rayleigh
// damping), K4 (velocity, momenta), K7 (density, mass fraction), periodic / outermost-layer
// copies and halo pack/unpack.  All are pure HBM streams (no reuse): one thread per point,
// i along threadIdx.x for coalescing.
#include <stdarg.h>

#include <atomic>

#include "stencil_math.cuh"

namespace tb200 {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace tb200

using namespace tb200;

extern "C" const char *tb200_last_error(void) { return g_err; }
extern "C" long long tb200_launch_count(void) { return g_launches.load(); }
extern "C" int tb200_version(void) { return 100; }
extern "C" int tb200_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return -TB200_ERR_CUDA;
  }
  return n;
}

// ---------------------------------------------------------------------------- K12
template <int OP>
__device__ __forceinline__ double ew(double a, double b, double c, double f) {
  if (OP == TB200_EW_COPY) return a;
  if (OP == TB200_EW_COPYCHANGE) return -a;
  if (OP == TB200_EW_ABS) return fabs(a);
  if (OP == TB200_EW_ADD) return a + b;
  if (OP == TB200_EW_ADDSUB) return a + b - c;
  if (OP == TB200_EW_CLIP) return a > 0.0 ? a : 0.0;
  if (OP == TB200_EW_FMA) return a + f * b;
  if (OP == TB200_EW_MUL) return a * b;
  if (OP == TB200_EW_SCALE) return f * a;
  if (OP == TB200_EW_SUB) return a - b;
  if (OP == TB200_EW_STS_RK2_0) return 0.5 * (a + b + f * c);
  if (OP == TB200_EW_STS_RK3WS_0) return (2.0 * a + b + f * c) / CDiv{3.0, 1.0 / 3.0};
  if (OP == TB200_EW_IADDSUB) return a + (b - c);
  if (OP == TB200_EW_ISCALE) return a * f;
  return 0.0;
}

template <int OP, int NIN>
static int run_ew(View out, View a, View b, View c, double f, const int32_t o[3],
                  const int32_t d[3], cudaStream_t st) {
  const int i0 = o[0], j0 = o[1], k0 = o[2];
  return launch_box("elementwise", d, st, [=] __device__(int i, int j, int k) {
    i += i0; j += j0; k += k0;
    const double va = a(i, j, k);
    const double vb = NIN >= 2 ? b(i, j, k) : 0.0;
    const double vc = NIN >= 3 ? c(i, j, k) : 0.0;
    out(i, j, k) = ew<OP>(va, vb, vc, f);
  });
}

extern "C" int tb200_elementwise(int op, tb200_field *out, const tb200_field *a,
                                 const tb200_field *b, const tb200_field *c, double f,
                                 const int32_t origin[3], const int32_t domain[3],
                                 void *stream) {
  View vo = view(out), va = view(a), vb = view(b), vc = view(c);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TB200_REQUIRE(box_inside(vo, origin, domain), "elementwise: out box outside storage");
  TB200_REQUIRE(box_inside(va, origin, domain), "elementwise: a box outside storage");
#define EW1(OP) case OP: return run_ew<OP, 1>(vo, va, vb, vc, f, origin, domain, st)
#define EW2(OP)                                                                        \
  case OP:                                                                             \
    TB200_REQUIRE(box_inside(vb, origin, domain), "elementwise: b box outside storage"); \
    return run_ew<OP, 2>(vo, va, vb, vc, f, origin, domain, st)
#define EW3(OP)                                                                        \
  case OP:                                                                             \
    TB200_REQUIRE(box_inside(vb, origin, domain), "elementwise: b box outside storage"); \
    TB200_REQUIRE(box_inside(vc, origin, domain), "elementwise: c box outside storage"); \
    return run_ew<OP, 3>(vo, va, vb, vc, f, origin, domain, st)
  switch (op) {
    EW1(TB200_EW_COPY);
    EW1(TB200_EW_COPYCHANGE);
    EW1(TB200_EW_ABS);
    EW2(TB200_EW_ADD);
    EW3(TB200_EW_ADDSUB);
    EW1(TB200_EW_CLIP);
    EW2(TB200_EW_FMA);
    EW2(TB200_EW_MUL);
    EW1(TB200_EW_SCALE);
    EW2(TB200_EW_SUB);
    EW3(TB200_EW_STS_RK2_0);
    EW3(TB200_EW_STS_RK3WS_0);
    EW3(TB200_EW_IADDSUB);
    EW1(TB200_EW_ISCALE);
    default:
      set_error("elementwise: unknown op %d", op);
      return TB200_ERR_ARG;
  }
}

// ---- coupler glue: one stage of a tendency stepper in one launch.
// The reference's DataArrayDictOperator.fma (src/tasmania/utils/xarrayx.py:L688-L740) runs the
// `fma` stencil once per stepped field; here every thread updates its point of all fields
// (up to TB200_FMA_MAX_FIELDS independent load pairs in flight), out_n = a_n + f * b_n.
namespace {
struct FmaFields {
  View out[TB200_FMA_MAX_FIELDS], a[TB200_FMA_MAX_FIELDS], b[TB200_FMA_MAX_FIELDS];
  int n;
};

// The generic box_kernel instantiation of this op needed 115 registers (eight predicated field
// slots per thread) and ran two 256-thread blocks per SM with two 8-byte loads per thread in
// flight: 1.0 TB/s, a quarter of the moist step and most of the diffusion dwarf's
// (profiles/README.md, round 2a).  Specialised on the number of fields instead; the vector
// variant gives a thread an aligned PAIR of columns (LDG.128 / STG.128, 2 N independent 16-byte
// loads in flight per thread).  out may alias a or b: a thread reads its own points first.
template <int N, bool VEC>
__global__ void __launch_bounds__(256) fma_fields_kernel(const FmaFields ff, double f, int i0, int j0,
                                                         int k0, int di, int dj, int dk) {
  const int j = j0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= j0 + dj) return;
  if (VEC) {
    const int c0 = (i0 & ~1) + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (c0 >= i0 + di) return;
    const bool m0 = c0 >= i0, m1 = c0 + 1 < i0 + di;
    for (int k = k0 + blockIdx.z; k < k0 + dk; k += gridDim.z) {
      double2 va[N], vb[N];
#pragma unroll
      for (int n = 0; n < N; ++n) {
        va[n] = *reinterpret_cast<const double2 *>(ff.a[n].p + (c0 + j * ff.a[n].s1 + k * ff.a[n].s2));
        vb[n] = *reinterpret_cast<const double2 *>(ff.b[n].p + (c0 + j * ff.b[n].s1 + k * ff.b[n].s2));
      }
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double *po = ff.out[n].p + (c0 + j * ff.out[n].s1 + k * ff.out[n].s2);
        const double r0 = va[n].x + f * vb[n].x, r1 = va[n].y + f * vb[n].y;
        if (m0 && m1)
          *reinterpret_cast<double2 *>(po) = make_double2(r0, r1);
        else if (m0)
          po[0] = r0;
        else if (m1)
          po[1] = r1;
      }
    }
  } else {
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i0 + di) return;
    for (int k = k0 + blockIdx.z; k < k0 + dk; k += gridDim.z) {
      double va[N], vb[N];
#pragma unroll
      for (int n = 0; n < N; ++n) {
        va[n] = ff.a[n](i, j, k);
        vb[n] = ff.b[n](i, j, k);
      }
#pragma unroll
      for (int n = 0; n < N; ++n) ff.out[n](i, j, k) = va[n] + f * vb[n];
    }
  }
}

// pairs may be read beyond the box (never beyond a row of the allocation: even pitches)
bool fma_vec_ok(const FmaFields &ff) {
  for (int n = 0; n < ff.n; ++n)
    for (const View *v : {&ff.out[n], &ff.a[n], &ff.b[n]})
      if (v->s0 != 1 || (v->s1 & 1) != 0 || (v->s2 & 1) != 0 || v->s1 < 2 ||
          (reinterpret_cast<uintptr_t>(v->p) & 15) != 0)
        return false;
  return true;
}

template <int N>
int launch_fma_fields(const FmaFields &ff, double f, const int32_t o[3], const int32_t d[3], cudaStream_t st) {
  if (d[0] <= 0 || d[1] <= 0 || d[2] <= 0) return TB200_OK;  // empty box: nothing to do
  const int gz = d[2] > 65535 ? 65535 : d[2];
  if (fma_vec_ok(ff)) {
    const int pairs = (o[0] + d[0] - (o[0] & ~1) + 1) / 2;
    dim3 block(pairs >= 128 ? 128 : (pairs > 32 ? 64 : 32), 1, 1);
    block.y = 256 / block.x;
    dim3 grid((pairs + block.x - 1) / block.x, (d[1] + block.y - 1) / block.y, gz);
    fma_fields_kernel<N, true><<<grid, block, 0, st>>>(ff, f, o[0], o[1], o[2], d[0], d[1], d[2]);
  } else {
    dim3 block(d[0] <= 32 ? 32 : 64, 1, 1);
    block.y = 256 / block.x;
    dim3 grid((d[0] + block.x - 1) / block.x, (d[1] + block.y - 1) / block.y, gz);
    fma_fields_kernel<N, false><<<grid, block, 0, st>>>(ff, f, o[0], o[1], o[2], d[0], d[1], d[2]);
  }
  return check_launch("fma_fields");
}
}  // namespace

extern "C" int tb200_fma_fields(int nfields, tb200_field *const *out,
                                const tb200_field *const *a, const tb200_field *const *b,
                                double f, const int32_t origin[3], const int32_t domain[3],
                                void *stream) {
  TB200_REQUIRE(out != nullptr && a != nullptr && b != nullptr, "fma_fields: NULL argument");
  TB200_REQUIRE(nfields >= 1 && nfields <= TB200_FMA_MAX_FIELDS, "fma_fields: 1..%d fields, got %d",
                TB200_FMA_MAX_FIELDS, nfields);
  FmaFields ff{};
  ff.n = nfields;
  for (int n = 0; n < nfields; ++n) {
    ff.out[n] = view(out[n]);
    ff.a[n] = view(a[n]);
    ff.b[n] = view(b[n]);
    TB200_REQUIRE(box_inside(ff.out[n], origin, domain), "fma_fields: out box outside storage %d", n);
    TB200_REQUIRE(box_inside(ff.a[n], origin, domain), "fma_fields: a box outside storage %d", n);
    TB200_REQUIRE(box_inside(ff.b[n], origin, domain), "fma_fields: b box outside storage %d", n);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (nfields) {
    case 1: return launch_fma_fields<1>(ff, f, origin, domain, st);
    case 2: return launch_fma_fields<2>(ff, f, origin, domain, st);
    case 3: return launch_fma_fields<3>(ff, f, origin, domain, st);
    case 4: return launch_fma_fields<4>(ff, f, origin, domain, st);
    case 5: return launch_fma_fields<5>(ff, f, origin, domain, st);
    case 6: return launch_fma_fields<6>(ff, f, origin, domain, st);
    case 7: return launch_fma_fields<7>(ff, f, origin, domain, st);
    default: return launch_fma_fields<8>(ff, f, origin, domain, st);
  }
}

// ---------------------------------------------------------------------------- K5
extern "C" int tb200_relax(const tb200_field *in_gamma, const tb200_field *in_phi,
                           const tb200_field *in_phi_ref, tb200_field *out_phi,
                           const int32_t origin[3], const int32_t domain[3], void *stream) {
  View g = view(in_gamma), r = view(in_phi_ref), o = view(out_phi);
  View p = in_phi ? view(in_phi) : o;  // irelax: in place
  TB200_REQUIRE(box_inside(g, origin, domain), "relax: gamma box outside storage");
  TB200_REQUIRE(box_inside(r, origin, domain), "relax: phi_ref box outside storage");
  TB200_REQUIRE(box_inside(o, origin, domain), "relax: out box outside storage");
  TB200_REQUIRE(box_inside(p, origin, domain), "relax: phi box outside storage");
  const int i0 = origin[0], j0 = origin[1], k0 = origin[2];
  return launch_box("relax", domain, static_cast<cudaStream_t>(stream),
                    [=] __device__(int i, int j, int k) {
                      i += i0; j += j0; k += k0;
                      o(i, j, k) = relax_point(g(i, j, k), p(i, j, k), r(i, j, k));
                    });
}

// Relaxed.enforce_raw in one launch over the boundary FRAME.  The reference relaxes every field
// with its own full-box `irelax` call (horizontal_boundary.py:L299-L344 -> relaxed.py:L119-L137)
// although gamma is zero -- and the point left untouched -- everywhere but on the nr outer rings
// (and the staggered extra row / column).  Here the host passes the largest box [i_lo, i_hi) x
// [j_lo, j_hi) on which gamma vanishes; the launch covers only its complement (west and east
// slabs over all rows, south and north slabs between them) for all fields at once: <10 % of the
// points, one launch instead of one per field.  Same point formula (relax_point): same bits.
namespace {
struct FrameFields {
  View phi[TB200_FMA_MAX_FIELDS], ref[TB200_FMA_MAX_FIELDS];
  int mi[TB200_FMA_MAX_FIELDS], mj[TB200_FMA_MAX_FIELDS], mk[TB200_FMA_MAX_FIELDS];
  int n;
};

__global__ void __launch_bounds__(256) relax_frame_kernel(const FrameFields ff, const View g, int i_lo,
                                                          int i_hi, int j_lo, int j_hi, int mi, int mj,
                                                          int mk) {
  const int w_e = mi - i_hi, w_c = i_hi - i_lo, h_n = mj - j_hi;
  const long long n_w = (long long)i_lo * mj, n_e = (long long)w_e * mj;
  const long long n_s = (long long)w_c * j_lo, n_n = (long long)w_c * h_n;
  const long long total = n_w + n_e + n_s + n_n;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    int i, j;
    if (p < n_w) {
      j = (int)(p / i_lo);
      i = (int)(p - (long long)j * i_lo);
    } else if (p < n_w + n_e) {
      const long long q = p - n_w;
      j = (int)(q / w_e);
      i = i_hi + (int)(q - (long long)j * w_e);
    } else if (p < n_w + n_e + n_s) {
      const long long q = p - n_w - n_e;
      j = (int)(q / w_c);
      i = i_lo + (int)(q - (long long)j * w_c);
    } else {
      const long long q = p - n_w - n_e - n_s;
      j = (int)(q / w_c);
      i = i_lo + (int)(q - (long long)j * w_c);
      j += j_hi;
    }
    const double gam = g(i, j, 0);
    if (gam == 0.0) continue;
    for (int k = blockIdx.y; k < mk; k += gridDim.y) {
#pragma unroll
      for (int n = 0; n < TB200_FMA_MAX_FIELDS; ++n) {
        if (n < ff.n && i < ff.mi[n] && j < ff.mj[n] && k < ff.mk[n])
          ff.phi[n](i, j, k) = relax_point(gam, ff.phi[n](i, j, k), ff.ref[n](i, j, k));
      }
    }
  }
}
}  // namespace

extern "C" int tb200_relax_frame(int nfields, tb200_field *const *phi,
                                 const tb200_field *const *phi_ref, const tb200_field *gamma,
                                 const int32_t *extents, const int32_t interior[4], void *stream) {
  TB200_REQUIRE(phi != nullptr && phi_ref != nullptr && extents != nullptr && interior != nullptr,
                "relax_frame: NULL argument");
  TB200_REQUIRE(nfields >= 1 && nfields <= TB200_FMA_MAX_FIELDS, "relax_frame: 1..%d fields, got %d",
                TB200_FMA_MAX_FIELDS, nfields);
  const View g = view(gamma);
  TB200_REQUIRE(g.ok(), "relax_frame: NULL gamma");
  FrameFields ff{};
  ff.n = nfields;
  int mi = 0, mj = 0, mk = 0;
  const int32_t o[3] = {0, 0, 0};
  for (int n = 0; n < nfields; ++n) {
    ff.phi[n] = view(phi[n]);
    ff.ref[n] = view(phi_ref[n]);
    const int32_t *d = extents + 3 * n;
    TB200_REQUIRE(box_inside(ff.phi[n], o, d), "relax_frame: extent outside storage of field %d", n);
    TB200_REQUIRE(box_inside(ff.ref[n], o, d), "relax_frame: extent outside reference field %d", n);
    ff.mi[n] = d[0];
    ff.mj[n] = d[1];
    ff.mk[n] = d[2];
    mi = d[0] > mi ? d[0] : mi;
    mj = d[1] > mj ? d[1] : mj;
    mk = d[2] > mk ? d[2] : mk;
  }
  TB200_REQUIRE(mi <= g.n0 && mj <= g.n1, "relax_frame: gamma smaller than the field extents");
  // clip the gamma-free box into the launch extent
  const int i_lo = interior[0] < 0 ? 0 : (interior[0] > mi ? mi : interior[0]);
  const int i_hi = interior[1] < i_lo ? i_lo : (interior[1] > mi ? mi : interior[1]);
  const int j_lo = interior[2] < 0 ? 0 : (interior[2] > mj ? mj : interior[2]);
  const int j_hi = interior[3] < j_lo ? j_lo : (interior[3] > mj ? mj : interior[3]);
  const long long total = (long long)(i_lo + mi - i_hi) * mj + (long long)(i_hi - i_lo) * (j_lo + mj - j_hi);
  if (total <= 0 || mk <= 0) return TB200_OK;
  const long long blocks = (total + 255) / 256;
  dim3 block(256, 1, 1);
  dim3 grid((unsigned)(blocks > 2048 ? 2048 : blocks), (unsigned)(mk > 65535 ? 65535 : mk), 1);
  relax_frame_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(ff, g, i_lo, i_hi, j_lo,
                                                                             j_hi, mi, mj, mk);
  return check_launch("relax_frame");
}

// Periodic.enforce_field: x wrap on rows [nb, my+nb), then y wrap over [0, mi)
extern "C" int tb200_periodic_enforce(tb200_field *field, int nx, int ny, int nb, int mx,
                                      int my, void *stream) {
  View f = view(field);
  TB200_REQUIRE(f.ok(), "periodic_enforce: NULL field");
  TB200_REQUIRE(nb > 0 && mx + 2 * nb <= f.n0 && my + 2 * nb <= f.n1,
                "periodic_enforce: field smaller than (mx+2nb, my+2nb)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sx = (mx == nx) ? 1 : 2, sy = (my == ny) ? 1 : 2;
  const int mi = mx + 2 * nb;
  // pass 1: both x-ghost slabs, 2*nb points along i
  {
    const int32_t d[3] = {2 * nb, my, f.n2};
    int rc = launch_box("periodic_x", d, st, [=] __device__(int i, int j, int k) {
      j += nb;
      if (i < nb) {
        f(i, j, k) = f(nx - 1 + i, j, k);
      } else {
        const int g = i - nb;
        f(mx + nb + g, j, k) = f(nb + sx + g, j, k);
      }
    });
    if (rc) return rc;
  }
  // pass 2: both y-ghost slabs (reads rows written by nobody in this pass)
  {
    const int32_t d[3] = {mi, 2 * nb, f.n2};
    return launch_box("periodic_y", d, st, [=] __device__(int i, int j, int k) {
      if (j < nb) {
        f(i, j, k) = f(i, ny - 1 + j, k);
      } else {
        const int g = j - nb;
        f(i, my + nb + g, k) = f(i, nb + sy + g, k);
      }
    });
  }
}

extern "C" int tb200_set_outermost_layers(tb200_field *field, const tb200_field *field_ref,
                                          int axis, int mi, int mj, void *stream) {
  View f = view(field), r = view(field_ref);
  TB200_REQUIRE(f.ok() && r.ok(), "set_outermost_layers: NULL field");
  TB200_REQUIRE(mi <= f.n0 && mj <= f.n1 && mi <= r.n0 && mj <= r.n1 && r.n2 >= f.n2,
                "set_outermost_layers: extents outside storage");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (axis == 0) {  // rows 0 and mi-1, j in [0, mj), every k  (relaxed.py:L161-L175)
    const int32_t d[3] = {2, mj, f.n2};
    return launch_box("outermost_x", d, st, [=] __device__(int i, int j, int k) {
      const int ii = i == 0 ? 0 : mi - 1;
      f(ii, j, k) = r(ii, j, k);
    });
  }
  const int32_t d[3] = {mi, 2, f.n2};  // relaxed.py:L177-L191
  return launch_box("outermost_y", d, st, [=] __device__(int i, int j, int k) {
    const int jj = j == 0 ? 0 : mj - 1;
    f(i, jj, k) = r(i, jj, k);
  });
}

// ---------------------------------------------------------------------------- K6
extern "C" int tb200_damping(const tb200_field *in_phi_now, const tb200_field *in_phi_new,
                             const tb200_field *in_phi_ref, const tb200_field *in_rmat,
                             tb200_field *out_phi, double dt, const int32_t origin[3],
                             const int32_t domain[3], void *stream) {
  View now = view(in_phi_now), nw = view(in_phi_new), ref = view(in_phi_ref);
  View rm = view(in_rmat), o = view(out_phi);
  TB200_REQUIRE(box_inside(now, origin, domain) && box_inside(nw, origin, domain) &&
                    box_inside(ref, origin, domain) && box_inside(rm, origin, domain) &&
                    box_inside(o, origin, domain),
                "damping: box outside storage");
  const int i0 = origin[0], j0 = origin[1], k0 = origin[2];
  return launch_box("damping", domain, static_cast<cudaStream_t>(stream),
                    [=] __device__(int i, int j, int k) {
                      i += i0; j += j0; k += k0;
                      o(i, j, k) =
                          damp_point(now(i, j, k), nw(i, j, k), ref(i, j, k), rm(i, j, k), dt);
                    });
}

// ---------------------------------------------------------------------------- Coriolis (8f-3)
// src/tasmania/isentropic/physics/coriolis.py:L166-L186: tnd_su = f sv, tnd_sv = -f su, each
// written or accumulated (set_output) over the box; one pass over both momenta.
extern "C" int tb200_coriolis(const tb200_field *in_su, const tb200_field *in_sv,
                              tb200_field *tnd_su, tb200_field *tnd_sv, double f, int ow_tnd_su,
                              int ow_tnd_sv, const int32_t origin[3], const int32_t domain[3],
                              void *stream) {
  View su = view(in_su), sv = view(in_sv), tu = view(tnd_su), tv = view(tnd_sv);
  TB200_REQUIRE(box_inside(su, origin, domain) && box_inside(sv, origin, domain) &&
                    box_inside(tu, origin, domain) && box_inside(tv, origin, domain),
                "coriolis: box outside storage");
  TB200_REQUIRE(tu.p != su.p && tu.p != sv.p && tv.p != su.p && tv.p != sv.p && tu.p != tv.p,
                "coriolis: tendencies must not alias the momenta or each other");
  const int i0 = origin[0], j0 = origin[1], k0 = origin[2];
  const bool owu = ow_tnd_su != 0, owv = ow_tnd_sv != 0;
  const double mf = -f;
  return launch_box("coriolis", domain, static_cast<cudaStream_t>(stream),
                    [=] __device__(int i, int j, int k) {
                      i += i0; j += j0; k += k0;
                      const double a = f * sv(i, j, k), b = mf * su(i, j, k);
                      tu(i, j, k) = owu ? a : tu(i, j, k) + a;
                      tv(i, j, k) = owv ? b : tv(i, j, k) + b;
                    });
}

// Coriolis forcing fused with the stage update of a tendency stepper (b200 only; the reference
// runs the stencil above into a tendency storage and then `fma` over the whole storage,
// framework/subclasses/tendency_steppers/*.py over utils/xarrayx.py:L688-L740):
//   out = base + factor * tendency,  tendency = f sv / -f su on [origin, origin + domain), 0 elsewhere
// over the `full` box (the storages' shape) in one pass: 4 reads + 2 writes per point instead of
// 2 + 2 (tendencies) and 4 + 2 (fma).  Same operations in the same order, hence the same bits.
extern "C" int tb200_coriolis_step(const tb200_field *in_su, const tb200_field *in_sv,
                                   const tb200_field *base_su, const tb200_field *base_sv,
                                   tb200_field *out_su, tb200_field *out_sv, double f, double factor,
                                   const int32_t origin[3], const int32_t domain[3],
                                   const int32_t full[3], void *stream) {
  View su = view(in_su), sv = view(in_sv), bu = view(base_su), bv = view(base_sv);
  View ou = view(out_su), ov = view(out_sv);
  const int32_t zero[3] = {0, 0, 0};
  TB200_REQUIRE(box_inside(su, origin, domain) && box_inside(sv, origin, domain),
                "coriolis_step: box outside an input storage");
  TB200_REQUIRE(box_inside(bu, zero, full) && box_inside(bv, zero, full) && box_inside(ou, zero, full) &&
                    box_inside(ov, zero, full),
                "coriolis_step: full box outside a base / output storage");
  TB200_REQUIRE(origin[0] >= 0 && origin[1] >= 0 && origin[2] >= 0 && origin[0] + domain[0] <= full[0] &&
                    origin[1] + domain[1] <= full[1] && origin[2] + domain[2] <= full[2],
                "coriolis_step: the tendency box must lie inside the full box");
  TB200_REQUIRE(ou.p != su.p && ou.p != sv.p && ov.p != su.p && ov.p != sv.p && ou.p != ov.p,
                "coriolis_step: outputs must not alias the inputs or each other");
  const int i0 = origin[0], j0 = origin[1], k0 = origin[2];
  const int i1 = i0 + domain[0], j1 = j0 + domain[1], k1 = k0 + domain[2];
  const double mf = -f;
  return launch_box("coriolis_step", full, static_cast<cudaStream_t>(stream),
                    [=] __device__(int i, int j, int k) {
                      double a = 0.0, b = 0.0;  // the tendency storages of the reference path hold 0 outside the box
                      if (i >= i0 && i < i1 && j >= j0 && j < j1 && k >= k0 && k < k1) {
                        a = f * sv(i, j, k);
                        b = mf * su(i, j, k);
                      }
                      ou(i, j, k) = bu(i, j, k) + factor * a;
                      ov(i, j, k) = bv(i, j, k) + factor * b;
                    });
}

// ---------------------------------------------------------------------------- K4
extern "C" int tb200_velocity(int axis, const tb200_field *in_d, const tb200_field *in_dw,
                              tb200_field *out_w, int staggering, const int32_t origin[3],
                              const int32_t domain[3], void *stream) {
  View d = view(in_d), dw = view(in_dw), w = view(out_w);
  const int hi = (staggering && axis == 0) ? 1 : 0, hj = (staggering && axis == 1) ? 1 : 0;
  TB200_REQUIRE(axis == 0 || axis == 1, "velocity: axis must be 0 or 1");
  TB200_REQUIRE(box_inside(d, origin, domain, hi, 0, hj, 0) &&
                    box_inside(dw, origin, domain, hi, 0, hj, 0) &&
                    box_inside(w, origin, domain),
                "velocity: box outside storage");
  const int i0 = origin[0], j0 = origin[1], k0 = origin[2];
  return launch_box("velocity", domain, static_cast<cudaStream_t>(stream),
                    [=] __device__(int i, int j, int k) {
                      i += i0; j += j0; k += k0;
                      if (staggering) {
                        w(i, j, k) = qdiv(dw(i - hi, j - hj, k) + dw(i, j, k),
                                          d(i - hi, j - hj, k) + d(i, j, k));
                      } else {
                        w(i, j, k) = qdiv(dw(i, j, k), d(i, j, k));
                      }
                    });
}

// Both velocity components of a stage output and their outermost faces in ONE pass
// (HorizontalVelocity.get_velocity_components, dwarfs/diagnostics.py:L219-L272, followed by
// Relaxed.set_outermost_layers_x / _y, relaxed.py:L161-L191): reads d, du, dv once (24 B/pt),
// writes u, v (16 B/pt).  The fused dry stage leaves the velocities of every stage to its successor
// (derive_uv_in); this kernel diagnoses them ONCE per time step, for the state the step returns.
// A lane owns an aligned pair of columns (LDG.128 / STG.128) and marches along j over a strip,
// carrying the previous row's d, dv; the column to the left comes from the neighbouring lane
// (lane 0: one scalar load); next row requested one row ahead, its lines pulled DRAM -> L2 four
// rows earlier.  Divisions through qdiv (common.cuh): same bits as `/`.
namespace {
constexpr int VXY_PF = 4;
__global__ void __launch_bounds__(128, 5) velocity_xy_kernel(View d, View du, View dv, View u, View v,
                                                          View ur, View vr, int nx, int ny, int lj) {
  const int lane = threadIdx.x & 31;
  const int w0 = 2 * (blockIdx.x * blockDim.x + (threadIdx.x & ~31));  // first column of the warp
  if (w0 >= nx) return;  // warp-uniform
  const int c0 = w0 + 2 * lane;
  const int pmax = (nx - 1) & ~1;           // last aligned pair that starts inside the row
  const int cm = min(c0, pmax);             // lanes beyond the row repeat its last pair (masked)
  const int cl = max(cm - 1, 0);            // lane 0's left neighbour
  const bool in0 = c0 < nx, in1 = c0 + 1 < nx;
  const int k = blockIdx.z;
  const int j0 = blockIdx.y * lj, j1 = min(j0 + lj, ny);
  const long long s1 = d.s1, pl = (long long)k * d.s2;
  auto ld2 = [](const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); };
  const double *pd = d.p + pl + cm, *pdu = du.p + pl + cm, *pdv = dv.p + pl + cm;
  double2 dp = make_double2(1.0, 1.0), dvp = make_double2(0.0, 0.0);  // row j - 1 (unused at j == 0)
  if (j0 > 0) {
    dp = ld2(pd + (j0 - 1) * s1);
    dvp = ld2(pdv + (j0 - 1) * s1);
  }
  double2 nd = ld2(pd + j0 * s1), ndu = ld2(pdu + j0 * s1), ndv = ld2(pdv + j0 * s1);
  double nl_d = 1.0, nl_du = 0.0;
  if (lane == 0 && c0 > 0) {
    nl_d = __ldg(d.p + pl + cl + j0 * s1);
    nl_du = __ldg(du.p + pl + cl + j0 * s1);
  }
  for (int j = j0; j < j1; ++j) {
    const double2 cd = nd, cdu = ndu, cdv = ndv;
    const double ld_ = nl_d, ldu_ = nl_du;
    const int jn = min(j + 1, ny - 1);  // the last iteration re-requests its own row (no overrun)
    nd = ld2(pd + jn * s1);
    ndu = ld2(pdu + jn * s1);
    ndv = ld2(pdv + jn * s1);
    if (lane == 0 && c0 > 0) {
      nl_d = __ldg(d.p + pl + cl + jn * s1);
      nl_du = __ldg(du.p + pl + cl + jn * s1);
    }
    if (j + VXY_PF < j1 && (lane & 7) == 0) {  // one request per 128-byte line and stream
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pd + (j + VXY_PF) * s1));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pdu + (j + VXY_PF) * s1));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pdv + (j + VXY_PF) * s1));
    }
    // left neighbours: the previous lane's second column
    double dl = __shfl_up_sync(0xffffffffu, cd.y, 1), dul = __shfl_up_sync(0xffffffffu, cdu.y, 1);
    if (lane == 0) { dl = ld_; dul = ldu_; }
    double u0, u1, v0, v1;
    if (c0 == 0) {
      u0 = ur.ok() ? ur.ld(0, j, k) : u.ld(0, j, k);
    } else {
      u0 = qdiv(dul + cdu.x, dl + cd.x);
    }
    u1 = qdiv(cdu.x + cdu.y, cd.x + cd.y);
    if (j == 0) {
      v0 = vr.ok() ? vr.ld(min(c0, nx - 1), 0, k) : v.ld(min(c0, nx - 1), 0, k);
      v1 = vr.ok() ? vr.ld(min(c0 + 1, nx - 1), 0, k) : v.ld(min(c0 + 1, nx - 1), 0, k);
    } else {
      v0 = qdiv(dvp.x + cdv.x, dp.x + cd.x);
      v1 = qdiv(dvp.y + cdv.y, dp.y + cd.y);
    }
    double *pu = u.p + pl + c0 + j * s1, *pv = v.p + pl + c0 + j * s1;
    if (in1) {
      *reinterpret_cast<double2 *>(pu) = make_double2(u0, u1);
      *reinterpret_cast<double2 *>(pv) = make_double2(v0, v1);
      if (c0 + 1 == nx - 1 && ur.ok()) pu[2] = ur.ld(nx, j, k);
      if (j == ny - 1 && vr.ok()) {
        pv[s1] = vr.ld(c0, ny, k);
        pv[s1 + 1] = vr.ld(c0 + 1, ny, k);
      }
    } else if (in0) {  // the last column of an odd-sized row
      pu[0] = u0;
      pv[0] = v0;
      if (ur.ok()) pu[1] = ur.ld(nx, j, k);
      if (j == ny - 1 && vr.ok()) pv[s1] = vr.ld(c0, ny, k);
    }
    dp = cd;
    dvp = cdv;
  }
}
}  // namespace

extern "C" int tb200_velocity_components(const tb200_field *in_d, const tb200_field *in_du,
                                         const tb200_field *in_dv, tb200_field *out_u,
                                         tb200_field *out_v, const tb200_field *u_ref,
                                         const tb200_field *v_ref, int nx, int ny, int nz,
                                         void *stream) {
  View d = view(in_d), du = view(in_du), dv = view(in_dv), u = view(out_u), v = view(out_v);
  View ur = view(u_ref), vr = view(v_ref);
  TB200_REQUIRE(nx >= 2 && ny >= 2 && nz >= 1, "velocity_components: need nx, ny >= 2, nz >= 1");
  const int32_t o[3] = {0, 0, 0}, m[3] = {nx, ny, nz}, mu[3] = {nx + 1, ny, nz}, mv[3] = {nx, ny + 1, nz};
  TB200_REQUIRE(box_inside(d, o, m) && box_inside(du, o, m) && box_inside(dv, o, m) &&
                    box_inside(u, o, mu) && box_inside(v, o, mv),
                "velocity_components: box outside storage");
  TB200_REQUIRE((!ur.ok() || box_inside(ur, o, mu)) && (!vr.ok() || box_inside(vr, o, mv)),
                "velocity_components: reference velocities too small");
  const View *all[] = {&d, &du, &dv, &u, &v};
  for (const View *f : all) {
    if (f->s0 != 1 || f->s1 != d.s1 || f->s2 != d.s2 || (f->s1 & 1) != 0 || (f->s2 & 1) != 0 ||
        (reinterpret_cast<uintptr_t>(f->p) & 15) != 0 || f->s1 < nx + 2) {
      set_error("velocity_components: the five fields must share one b200 storage geometry (unit i-stride, "
                "equal even pitches with two spare columns, 16-byte aligned)");
      return TB200_ERR_LAYOUT;
    }
  }
  const int warps_x = (nx + 63) / 64;
  const int wpb = warps_x >= 4 ? 4 : warps_x;  // warps per block, side by side
  const int gx = (warps_x + wpb - 1) / wpb;
  int lj = 64;
  if ((long long)gx * ((ny + 63) / 64) * nz < 148 * 4) lj = 8;
  dim3 grid(gx, (ny + lj - 1) / lj, nz);
  velocity_xy_kernel<<<grid, 32 * wpb, 0, static_cast<cudaStream_t>(stream)>>>(d, du, dv, u, v, ur, vr, nx,
                                                                                 ny, lj);
  return check_launch("velocity_components");
}

extern "C" int tb200_momenta(const tb200_field *in_d, const tb200_field *in_u,
                             const tb200_field *in_v, tb200_field *out_du,
                             tb200_field *out_dv, int staggering, const int32_t origin[3],
                             const int32_t domain[3], void *stream) {
  View d = view(in_d), u = view(in_u), v = view(in_v), du = view(out_du), dv = view(out_dv);
  const int h = staggering ? 1 : 0;
  TB200_REQUIRE(box_inside(d, origin, domain) && box_inside(u, origin, domain, 0, h) &&
                    box_inside(v, origin, domain, 0, 0, 0, h) &&
                    box_inside(du, origin, domain) && box_inside(dv, origin, domain),
                "momenta: box outside storage");
  const int i0 = origin[0], j0 = origin[1], k0 = origin[2];
  return launch_box("momenta", domain, static_cast<cudaStream_t>(stream),
                    [=] __device__(int i, int j, int k) {
                      i += i0; j += j0; k += k0;
                      if (staggering) {
                        du(i, j, k) = 0.5 * d(i, j, k) * (u(i, j, k) + u(i + 1, j, k));
                        dv(i, j, k) = 0.5 * d(i, j, k) * (v(i, j, k) + v(i, j + 1, k));
                      } else {
                        du(i, j, k) = d(i, j, k) * u(i, j, k);
                        dv(i, j, k) = d(i, j, k) * v(i, j, k);
                      }
                    });
}

// ---------------------------------------------------------------------------- K7
extern "C" int tb200_density(const tb200_field *in_d, const tb200_field *in_q,
                             tb200_field *out_dq, int clipping, const int32_t origin[3],
                             const int32_t domain[3], void *stream) {
  View d = view(in_d), q = view(in_q), dq = view(out_dq);
  TB200_REQUIRE(box_inside(d, origin, domain) && box_inside(q, origin, domain) &&
                    box_inside(dq, origin, domain),
                "density: box outside storage");
  const int i0 = origin[0], j0 = origin[1], k0 = origin[2];
  return launch_box("density", domain, static_cast<cudaStream_t>(stream),
                    [=] __device__(int i, int j, int k) {
                      i += i0; j += j0; k += k0;
                      const double x = d(i, j, k) * q(i, j, k);
                      dq(i, j, k) = clipping ? (x > 0.0 ? x : 0.0) : x;
                    });
}

extern "C" int tb200_mass_fraction(const tb200_field *in_d, const tb200_field *in_dq,
                                   tb200_field *out_q, int clipping, const int32_t origin[3],
                                   const int32_t domain[3], void *stream) {
  View d = view(in_d), dq = view(in_dq), q = view(out_q);
  TB200_REQUIRE(box_inside(d, origin, domain) && box_inside(dq, origin, domain) &&
                    box_inside(q, origin, domain),
                "mass_fraction: box outside storage");
  const int i0 = origin[0], j0 = origin[1], k0 = origin[2];
  return launch_box("mass_fraction", domain, static_cast<cudaStream_t>(stream),
                    [=] __device__(int i, int j, int k) {
                      i += i0; j += j0; k += k0;
                      const double x = qdiv(dq(i, j, k), d(i, j, k));
                      q(i, j, k) = clipping ? (x > 0.0 ? x : 0.0) : x;
                    });
}

// ---------------------------------------------------------------------------- halo boxes
extern "C" int tb200_pack_box(const tb200_field *field, double *buffer,
                              const int32_t origin[3], const int32_t domain[3], void *stream) {
  View f = view(field);
  TB200_REQUIRE(box_inside(f, origin, domain) && buffer, "pack_box: bad arguments");
  const int i0 = origin[0], j0 = origin[1], k0 = origin[2];
  const long long di = domain[0], dj = domain[1];
  return launch_box("pack_box", domain, static_cast<cudaStream_t>(stream),
                    [=] __device__(int i, int j, int k) {
                      buffer[i + di * (j + dj * k)] = f(i + i0, j + j0, k + k0);
                    });
}

extern "C" int tb200_unpack_box(tb200_field *field, const double *buffer,
                                const int32_t origin[3], const int32_t domain[3],
                                void *stream) {
  View f = view(field);
  TB200_REQUIRE(box_inside(f, origin, domain) && buffer, "unpack_box: bad arguments");
  const int i0 = origin[0], j0 = origin[1], k0 = origin[2];
  const long long di = domain[0], dj = domain[1];
  return launch_box("unpack_box", domain, static_cast<cudaStream_t>(stream),
                    [=] __device__(int i, int j, int k) {
                      f(i + i0, j + j0, k + k0) = buffer[i + di * (j + dj * k)];
                    });
}

// ---- self-test of qdiv (common.cuh) against the compiler's IEEE division: `count` operand pairs
// from a counter-based generator, in classes that cover the fast path, its guard and the known
// hard cases of Newton-Raphson division (divisor mantissa all ones, quotients next to a rounding
// boundary, numerator zero / tiny / huge).  *mismatches = pairs whose results differ in any bit
// (+0 and -0 count as equal: a zero numerator keeps the fast path, which returns +0 for -0 / b).
namespace {
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double make_fp(unsigned long long mant, int exp2, bool neg) {
  const unsigned long long bits = ((unsigned long long)neg << 63) | ((unsigned long long)(exp2 + 1023) << 52) |
                                  (mant & 0xfffffffffffffull);
  return __longlong_as_double((long long)bits);
}
__global__ void __launch_bounds__(256) qdiv_selftest_kernel(unsigned long long count, unsigned long long seed,
                                                            unsigned long long *mismatches) {
  unsigned long long bad = 0;
  for (unsigned long long n = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; n < count;
       n += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long r0 = mix64(seed + 3 * n), r1 = mix64(seed + 3 * n + 1), r2 = mix64(seed + 3 * n + 2);
    const int cls = (int)(r2 & 15);
    // defaults: the data range of the velocity diagnosis (b a sum of densities, a a sum of momenta)
    int ea = (int)((r2 >> 8) % 41) - 20, eb = (int)((r2 >> 16) % 21) - 4;
    unsigned long long ma = r0, mb = r1;
    bool neg = (r2 >> 4) & 1;
    if (cls == 1) ea = (int)((r2 >> 8) % 1200) - 760;          // wide numerators, across the guard
    if (cls == 2) eb = (int)((r2 >> 16) % 300) - 150;          // wide divisors, across the guard
    if (cls == 3) mb = 0xfffffffffffffull - ((r1 >> 20) & 3);  // divisor mantissa (almost) all ones
    if (cls == 4) mb = (r1 >> 20) & 7;                         // divisor next to a power of two
    if (cls == 5) ma = mb + ((r0 >> 20) & 7) - 3;              // quotient next to 1
    if (cls == 6) ma = 0xfffffffffffffull - ((r0 >> 20) & 3);
    if (cls == 7) ma = mb = 0;                                 // powers of two
    double a = make_fp(ma, ea, neg), b = make_fp(mb, eb, false);
    if (cls == 8) a = neg ? -0.0 : 0.0;
    if (cls == 9) a = __longlong_as_double((long long)(r0 & 0xfffffffffffffull));  // subnormal numerator
    if (cls == 10) b = (double)(1 + (r1 & 0xffff)) * 0.5;      // small half-integers
    if (cls == 11) { a = (double)(long long)(r0 >> 34) * 0.25; b = (double)(1 + (r1 >> 44)); }
    const double q = tb200::qdiv(a, b), want = a / b;
    const bool same = __double_as_longlong(q) == __double_as_longlong(want) || (q == 0.0 && want == 0.0);
    bad += same ? 0ull : 1ull;
  }
  if (bad) atomicAdd(mismatches, bad);
}
}  // namespace

extern "C" int tb200_selftest_division(uint64_t count, uint64_t seed, uint64_t *mismatches, void *stream) {
  TB200_REQUIRE(mismatches != nullptr, "selftest_division: NULL result pointer (device memory, zeroed)");
  if (count == 0) return TB200_OK;
  qdiv_selftest_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      count, seed, reinterpret_cast<unsigned long long *>(mismatches));
  return check_launch("selftest_division");
}
