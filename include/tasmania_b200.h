/* tasmania_b200.h -- C ABI of libtasmania_b200.so
 *
 * B200-native (sm_100a, fp64) replacement for the stencils on tasmania's hot path
 * (SURVEY.md section 8a, K1..K12).  tasmania is pure Python: the "FFI" these entry points
 * replace is the call of a compiled stencil object,
 *
 *     stencil(**arrays, **scalars, origin=(i0,j0,k0), domain=(di,dj,dk), ...)
 *
 * returned by `StencilFactory.compile_stencil(name)`
 * (reference: src/tasmania/framework/stencil.py:L273-L284, the numpy "compiler" being
 * src/tasmania/framework/subclasses/stencil_compilers.py:L91-L99).  Each function below
 * cites the reference stencil definition it stands in for; argument names follow the
 * reference's keyword names.  The Python side (tasmania_b200/lib.py) binds them with ctypes
 * and fills `tb200_field` from `__cuda_array_interface__`.
 *
 * Conventions
 *  - every field is fp64 device memory, addressed with explicit element strides, so sliced
 *    (non-contiguous) views are fine; the library never owns, frees or retains field memory;
 *  - `origin`/`domain` have the reference's meaning (first point / number of points);
 *  - optional fields are passed as NULL;
 *  - kernels are enqueued on `stream` (a cudaStream_t; NULL = legacy default stream) and the
 *    call returns immediately: no device synchronisation inside the library;
 *  - return value 0 = success, otherwise an error code; `tb200_last_error()` gives the
 *    thread-local message.  There is no CPU fallback: a call without a usable CUDA device
 *    fails with TB200_ERR_CUDA.
 */
#ifndef TASMANIA_B200_H
#define TASMANIA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tb200_field {
  void *ptr;         /* device pointer to element (0,0,0), fp64 */
  int64_t shape[3];  /* logical extents (ni, nj, nk) */
  int64_t stride[3]; /* strides in ELEMENTS */
} tb200_field;

enum {
  TB200_OK = 0,
  TB200_ERR_ARG = 1,   /* bad argument (NULL field, box outside the storage, bad enum) */
  TB200_ERR_CUDA = 2,  /* CUDA runtime error (message holds cudaGetErrorString) */
  TB200_ERR_LAYOUT = 3 /* fused kernels need unit stride along i */
};

/* horizontal flux schemes: src/tasmania/isentropic/dynamics/subclasses/
 * minimal_horizontal_fluxes/{upwind,centered,third_order_upwind,fifth_order_upwind}.py */
enum {
  TB200_FLUX_UPWIND = 0,
  TB200_FLUX_CENTERED = 1,
  TB200_FLUX_THIRD_ORDER_UPWIND = 2,
  TB200_FLUX_FIFTH_ORDER_UPWIND = 3
};

/* element-wise stencils: src/tasmania/framework/subclasses/stencil_definitions/
 * copy.py:L30-L41, math.py:L32-L124, algorithms.py:L60-L69 */
enum {
  TB200_EW_COPY = 0,        /* out = a                         */
  TB200_EW_COPYCHANGE = 1,  /* out = -a                        */
  TB200_EW_ABS = 2,         /* out = |a|                       */
  TB200_EW_ADD = 3,         /* out = a + b                     */
  TB200_EW_ADDSUB = 4,      /* out = a + b - c                 */
  TB200_EW_CLIP = 5,        /* out = a > 0 ? a : 0             */
  TB200_EW_FMA = 6,         /* out = a + f * b                 */
  TB200_EW_MUL = 7,         /* out = a * b                     */
  TB200_EW_SCALE = 8,       /* out = f * a                     */
  TB200_EW_SUB = 9,         /* out = a - b                     */
  TB200_EW_STS_RK2_0 = 10,  /* out = 0.5 * (a + b + f * c)     */
  TB200_EW_STS_RK3WS_0 = 11,/* out = (2 a + b + f * c) / 3     */
  TB200_EW_IADDSUB = 12,    /* out = a + (b - c)  (in-place `+=` form, math.py:L84-L88) */
  TB200_EW_ISCALE = 13      /* out = a * f        (in-place `*=` form, math.py:L103-L106) */
};

const char *tb200_last_error(void);
int tb200_version(void);
/* number of visible CUDA devices, or a negative error code */
int tb200_device_count(void);
/* kernels this library has launched since it was loaded (every successful launch counts;
 * bench.py reports the difference over its timed region as `gpu_launches`) */
long long tb200_launch_count(void);

/* ---- scratch memory behind an explicit handle ------------------
 * The library never owns field memory.  What the fused kernels need beyond the fields (the hand-off
 * arrays of tb200_isentropic_stage_dry / _moist: scratch_exn, scratch_mtg, scratch_s) is requested
 * from a context held by the host-side backend object: tb200_ctx_scratch returns `count` zero-filled
 * fields of logical shape `shape` in the b200 storage layout (unit i-stride, rows padded to 16
 * doubles), the same ones on every call with that shape, valid until tb200_ctx_destroy.
 * Reference counterpart: the temporaries its objects allocate for themselves, e.g. `_mtg_new` of
 * the prognostic (isentropic/dynamics/subclasses/prognostics/rk3ws_si.py:L241) and the sq* / *_ref
 * storages of the dycore (isentropic/dynamics/dycore.py:L321-L345). */
typedef struct tb200_ctx tb200_ctx;
int tb200_ctx_create(tb200_ctx **ctx);
int tb200_ctx_destroy(tb200_ctx *ctx);
int tb200_ctx_scratch(tb200_ctx *ctx, const int64_t shape[3], int count, tb200_field *fields);

/* ---- K12 element-wise ------------------------------------------------------------- */
int tb200_elementwise(int op, tb200_field *out, const tb200_field *a, const tb200_field *b,
                      const tb200_field *c, double f, const int32_t origin[3],
                      const int32_t domain[3], void *stream);

/* Coupler glue: out[n] = a[n] + f * b[n] for n < nfields in ONE launch --
 * one stage of a tendency stepper over all the fields it steps.  Replaces the per-field `fma`
 * calls of DataArrayDictOperator.fma (src/tasmania/utils/xarrayx.py:L688-L740, stencil
 * src/tasmania/framework/subclasses/stencil_definitions/math.py:L59-L63) issued by
 * src/tasmania/framework/subclasses/tendency_steppers/{forward_euler,rk2,rk3ws}.py. */
#define TB200_FMA_MAX_FIELDS 8
int tb200_fma_fields(int nfields, tb200_field *const *out, const tb200_field *const *a,
                     const tb200_field *const *b, double f, const int32_t origin[3],
                     const int32_t domain[3], void *stream);

/* ---- K5 lateral boundaries --------------------------------------------------------- */
/* algorithms.py:L32-L43 (irelax; pass in_phi == NULL) and L46-L57 (relax) */
int tb200_relax(const tb200_field *in_gamma, const tb200_field *in_phi,
                const tb200_field *in_phi_ref, tb200_field *out_phi, const int32_t origin[3],
                const int32_t domain[3], void *stream);
/* Relaxed.enforce_raw (horizontal_boundary.py:L299-L344 over relaxed.py:L119-L137) for up to
 * TB200_FMA_MAX_FIELDS fields in one launch restricted to the frame where gamma != 0:
 * extents = nfields x (mi, mj, mk) of each field (staggered fields are one point larger),
 * interior = (i_lo, i_hi, j_lo, j_hi), a box on which gamma is zero (points inside are skipped). */
int tb200_relax_frame(int nfields, tb200_field *const *phi, const tb200_field *const *phi_ref,
                      const tb200_field *gamma, const int32_t *extents, const int32_t interior[4],
                      void *stream);

/* Periodic.enforce_field, src/tasmania/domain/subclasses/horizontal_boundaries/
 * periodic.py:L98-L122; nx, ny = physical sizes, mx, my = nx|nx+1, ny|ny+1 (staggering) */
int tb200_periodic_enforce(tb200_field *field, int nx, int ny, int nb, int mx, int my,
                           void *stream);
/* Relaxed.set_outermost_layers_x / _y, relaxed.py:L161-L191; axis 0 -> x, 1 -> y */
int tb200_set_outermost_layers(tb200_field *field, const tb200_field *field_ref, int axis,
                               int mi, int mj, void *stream);

/* ---- K6 Rayleigh damping: src/tasmania/dwarfs/subclasses/vertical_dampers/
 * rayleigh.py:L90-L109 */
int tb200_damping(const tb200_field *in_phi_now, const tb200_field *in_phi_new,
                  const tb200_field *in_phi_ref, const tb200_field *in_rmat,
                  tb200_field *out_phi, double dt, const int32_t origin[3],
                  const int32_t domain[3], void *stream);

/* ---- K4 / K7 diagnostics: src/tasmania/dwarfs/diagnostics.py:L175-L272, L400-L450 --- */
int tb200_velocity(int axis, const tb200_field *in_d, const tb200_field *in_dw,
                   tb200_field *out_w, int staggering, const int32_t origin[3],
                   const int32_t domain[3], void *stream);
/* Both staggered velocity components of a state and their outermost faces in one pass over
 * [0, nx) x [0, ny) x [0, nz): HorizontalVelocity.get_velocity_components
 * (src/tasmania/dwarfs/diagnostics.py:L219-L272, staggering=True) followed by
 * Relaxed.set_outermost_layers_x / _y (src/tasmania/domain/subclasses/horizontal_boundaries/
 * relaxed.py:L161-L191): u(i) = (du(i-1) + du(i)) / (d(i-1) + d(i)) for 0 < i < nx, u(0), u(nx)
 * from u_ref; likewise v along j.  u_ref / v_ref may be NULL (outermost faces untouched).
 * The five fields must share one b200 storage geometry (else TB200_ERR_LAYOUT: use
 * tb200_velocity + tb200_set_outermost_layers). */
int tb200_velocity_components(const tb200_field *in_d, const tb200_field *in_du,
                              const tb200_field *in_dv, tb200_field *out_u, tb200_field *out_v,
                              const tb200_field *u_ref, const tb200_field *v_ref, int nx, int ny,
                              int nz, void *stream);
int tb200_momenta(const tb200_field *in_d, const tb200_field *in_u, const tb200_field *in_v,
                  tb200_field *out_du, tb200_field *out_dv, int staggering,
                  const int32_t origin[3], const int32_t domain[3], void *stream);
int tb200_density(const tb200_field *in_d, const tb200_field *in_q, tb200_field *out_dq,
                  int clipping, const int32_t origin[3], const int32_t domain[3],
                  void *stream);
int tb200_mass_fraction(const tb200_field *in_d, const tb200_field *in_dq, tb200_field *out_q,
                        int clipping, const int32_t origin[3], const int32_t domain[3],
                        void *stream);

/* ---- K8 diffusion: src/tasmania/dwarfs/subclasses/horizontal_diffusers/
 * second_order.py:L92-L106, fourth_order.py:L92-L124 (+ set_output, generics.py:L38-L40).
 * order = 2 | 4.  k runs over [origin[2], origin[2]+domain[2]); the reference's numpy
 * definition ignores the k-range and processes every level of the storage -- the Python
 * stencil wrapper passes the full k extent to reproduce that. */
int tb200_diffusion(int order, const tb200_field *in_phi, const tb200_field *in_gamma,
                    tb200_field *out_phi, double dx, double dy, int ow_out_phi,
                    const int32_t origin[3], const int32_t domain[3], void *stream);

/* ---- K9 smoothing: src/tasmania/dwarfs/subclasses/horizontal_smoothers/
 * first_order.py:L113-L126, second_order.py:L113-L139, third_order.py:L113-L150.
 * order = 1 | 2 | 3.  With rim_copy != 0 the four `copy` launches of
 * HorizontalSmoothing.__call__ (first_order.py:L77-L110) are fused in: points of the
 * (2*origin[0]+domain[0], 2*origin[1]+domain[1]) box -- the smoother's shape -- that lie
 * outside [origin, origin+domain) are copied from in_phi. */
int tb200_smoothing(int order, const tb200_field *in_phi, const tb200_field *in_gamma,
                    tb200_field *out_phi, int rim_copy, const int32_t origin[3],
                    const int32_t domain[3], void *stream);

/* ---- K8 / K9, one-dimensional variants: SecondOrder1DX / 1DY (second_order.py:L150-L219,
 * L261-L330), FourthOrder1DX / 1DY (fourth_order.py:L196-L277, L333-L414) of the diffusers and
 * FirstOrder1DX / 1DY (first_order.py:L141-L218, L233-L310), SecondOrder1DX / 1DY
 * (second_order.py:L163-L247, L267-L351), ThirdOrder1DX / 1DY (third_order.py:L175-L263,
 * L285-L373) of the smoothers.  axis = 0 (x) | 1 (y) is the axis the stencil runs along; h is
 * the grid spacing along it (the reference's definitions receive dx and dy and use one).
 * Same k-range and rim_copy conventions as the two-dimensional entry points (a 1-D smoother
 * copies the two rim slabs across its axis: first_order.py:L187-L204). */
int tb200_diffusion_1d(int order, int axis, const tb200_field *in_phi,
                       const tb200_field *in_gamma, tb200_field *out_phi, double h,
                       int ow_out_phi, const int32_t origin[3], const int32_t domain[3],
                       void *stream);
int tb200_smoothing_1d(int order, int axis, const tb200_field *in_phi,
                       const tb200_field *in_gamma, tb200_field *out_phi, int rim_copy,
                       const int32_t origin[3], const int32_t domain[3], void *stream);

/* ---- the class-less `diffusion` stencil (a hyperdiffusion filter):
 * src/tasmania/framework/subclasses/stencil_definitions/diffusion.py:L31-L55.
 * out = phi + alpha * div(grad(lap(lap(phi)))) on [origin, origin+domain); reads a halo of 3. */
int tb200_hyperdiffusion(const tb200_field *in_phi, tb200_field *out_phi, double alpha,
                         const int32_t origin[3], const int32_t domain[3], void *stream);

/* ---- K1 / K2 isentropic prognostic step: src/tasmania/isentropic/dynamics/subclasses/
 * prognostics/utils.py:L43-L134 and L137-L204.  The moist tracers are passed as arrays of
 * three field pointers (qv, qc, qr order); NULL arrays = dry. */
int tb200_step_forward_euler(int flux_scheme, const tb200_field *s_now,
                             const tb200_field *s_int, tb200_field *s_new,
                             const tb200_field *u_int, const tb200_field *v_int,
                             const tb200_field *s_tnd, const tb200_field *const *sq_now,
                             const tb200_field *const *sq_int, tb200_field *const *sq_new,
                             const tb200_field *const *q_tnd, double dt, double dx, double dy,
                             const int32_t origin[3], const int32_t domain[3], void *stream);
int tb200_step_forward_euler_momentum(
    int flux_scheme, const tb200_field *s_now, const tb200_field *s_new,
    const tb200_field *u_int, const tb200_field *v_int, const tb200_field *su_now,
    const tb200_field *su_int, tb200_field *su_new, const tb200_field *sv_now,
    const tb200_field *sv_int, tb200_field *sv_new, const tb200_field *mtg_now,
    const tb200_field *mtg_new, const tb200_field *su_tnd, const tb200_field *sv_tnd,
    double dt, double dx, double dy, double eps, const int32_t origin[3],
    const int32_t domain[3], void *stream);

/* ---- K3 column scans: src/tasmania/isentropic/dynamics/diagnostics.py
 * montgomery L408-L438, diagnostic_variables L319-L360, height L472-L503,
 * density_and_temperature L540-L570.  constants = {pref, rd, g, cp}.
 * in_theta may be NULL for montgomery (theta_s is passed instead). */
int tb200_montgomery(const tb200_field *in_hs, const tb200_field *in_s,
                     tb200_field *inout_mtg, double dz, double pt, double theta_s,
                     const double constants[4], const int32_t origin[3],
                     const int32_t domain[3], void *stream);
int tb200_diagnostic_variables(const tb200_field *in_theta, const tb200_field *in_hs,
                               const tb200_field *in_s, tb200_field *inout_p,
                               tb200_field *out_exn, tb200_field *inout_mtg,
                               tb200_field *inout_h, double dz, double pt,
                               const double constants[4], const int32_t origin[3],
                               const int32_t domain[3], void *stream);
int tb200_height(const tb200_field *in_theta, const tb200_field *in_hs,
                 const tb200_field *in_s, tb200_field *inout_h, double dz, double pt,
                 const double constants[4], const int32_t origin[3],
                 const int32_t domain[3], void *stream);
int tb200_density_and_temperature(const tb200_field *in_theta, const tb200_field *in_s,
                                  const tb200_field *in_exn, const tb200_field *in_h,
                                  tb200_field *out_rho, tb200_field *out_t, double cp,
                                  const int32_t origin[3], const int32_t domain[3],
                                  void *stream);

/* ---- K10 Burgers: src/tasmania/burgers/dynamics/stepper.py:L188-L227 with the advection
 * subroutines of burgers/dynamics/subclasses/advection/{first..sixth}_order.py;
 * advection_order = 1..6 */
int tb200_burgers_forward_euler(int advection_order, const tb200_field *in_u,
                                const tb200_field *in_v, const tb200_field *in_u_tmp,
                                const tb200_field *in_v_tmp, tb200_field *out_u,
                                tb200_field *out_v, const tb200_field *in_u_tnd,
                                const tb200_field *in_v_tnd, double dt, double dx, double dy,
                                const int32_t origin[3], const int32_t domain[3],
                                void *stream);

/* ---- K11 Kessler microphysics: src/tasmania/physics/microphysics/kessler.py
 * kessler L307-L376, saturation (diagnostic) L661-L714, saturation (prognostic) L981-L1032,
 * fall_velocity L1183-L1203, sedimentation L1339-L1370 (+ sedimentation_fluxes/{first,second}
 * _order.py), accumulated_precipitation microphysics/utils.py:L283-L305.
 * `flags`: the compile-time externals and the `ow_*` (set_output overwrite) switches. */
enum {
  TB200_KESSLER_P_ON_INTERFACES = 1,  /* air_pressure_on_interface_levels: p, exn averaged k, k+1 */
  TB200_KESSLER_RAIN_EVAPORATION = 2,
  TB200_KESSLER_OW_QC = 4,
  TB200_KESSLER_OW_QR = 8,
  TB200_KESSLER_OW_QV = 16,
  TB200_KESSLER_OW_THETA = 32
};
int tb200_kessler(const tb200_field *in_rho, const tb200_field *in_p, const tb200_field *in_t,
                  const tb200_field *in_exn, const tb200_field *in_qc, const tb200_field *in_qr,
                  const tb200_field *in_qv, tb200_field *out_qc_tnd, tb200_field *out_qr_tnd,
                  tb200_field *out_qv_tnd, tb200_field *out_theta_tnd, double a, double k1,
                  double k2, double beta, double lhvw, uint32_t flags, const int32_t origin[3],
                  const int32_t domain[3], void *stream);
int tb200_saturation_diagnostic(const tb200_field *in_p, const tb200_field *in_t,
                                const tb200_field *in_exn, const tb200_field *in_qv,
                                const tb200_field *in_qc, tb200_field *out_qv,
                                tb200_field *out_qc, tb200_field *out_t, tb200_field *tnd_theta,
                                double dt, double beta, double lhvw, double cp, double rv,
                                uint32_t flags, const int32_t origin[3],
                                const int32_t domain[3], void *stream);
int tb200_saturation_prognostic(const tb200_field *in_p, const tb200_field *in_t,
                                const tb200_field *in_exn, const tb200_field *in_qv,
                                const tb200_field *in_qc, tb200_field *tnd_qv,
                                tb200_field *tnd_qc, tb200_field *tnd_theta, double sr,
                                double beta, double lhvw, double cp, double rv, uint32_t flags,
                                const int32_t origin[3], const int32_t domain[3], void *stream);
int tb200_fall_velocity(const tb200_field *in_rho, const tb200_field *in_rho_s,
                        const tb200_field *in_qr, tb200_field *out_vt, const int32_t origin[3],
                        const int32_t domain[3], void *stream);
/* order = 1 | 2: first / second order upwind sedimentation flux (sflux_extent = order) */
int tb200_sedimentation(int order, const tb200_field *in_rho, const tb200_field *in_h,
                        const tb200_field *in_qr, const tb200_field *in_vt,
                        tb200_field *out_tnd_qr, int ow_out_tnd_qr, const int32_t origin[3],
                        const int32_t domain[3], void *stream);
int tb200_accumulated_precipitation(const tb200_field *in_rho, const tb200_field *in_qr,
                                    const tb200_field *in_vt, const tb200_field *in_accprec,
                                    tb200_field *out_prec, tb200_field *out_accprec, double dt,
                                    double rhow, const int32_t origin[3],
                                    const int32_t domain[3], void *stream);

/* ---- Coriolis forcing of the momenta:
 * src/tasmania/isentropic/physics/coriolis.py:L166-L186  tnd_su (+)= f sv, tnd_sv (+)= -f su */
int tb200_coriolis(const tb200_field *in_su, const tb200_field *in_sv, tb200_field *tnd_su,
                   tb200_field *tnd_sv, double f, int ow_tnd_su, int ow_tnd_sv,
                   const int32_t origin[3], const int32_t domain[3], void *stream);

/* ---- Smagorinsky horizontal turbulence:
 * src/tasmania/physics/turbulence.py:L165-L229 (in_s NULL: in_a = u, in_b = v, tendencies of u, v)
 * src/tasmania/isentropic/physics/turbulence.py:L99-L125 (in_s, in_a = su, in_b = sv: u = su / s,
 * v = sv / s, tendencies of su, sv).  The box needs two more points on every horizontal side. */
int tb200_smagorinsky(const tb200_field *in_s, const tb200_field *in_a, const tb200_field *in_b,
                      tb200_field *out_a_tnd, tb200_field *out_b_tnd, double dx, double dy,
                      double cs, int ow_out_a_tnd, int ow_out_b_tnd, const int32_t origin[3],
                      const int32_t domain[3], void *stream);

/* ---- vertical advection: IsentropicVerticalAdvection._stencil
 * src/tasmania/isentropic/physics/vertical_advection.py:L271-L386 with the minimal vertical flux
 * schemes of src/tasmania/isentropic/dynamics/subclasses/minimal_vertical_fluxes/ (flux_scheme =
 * TB200_FLUX_*).  staggered_w: in_w lives on the interface levels (externals["staggering"]), else
 * it is averaged onto them.  in_q* / out_q* NULL = dry.  overwrite_flags: bit f = ow_out_<field f>
 * for the fields (s, su, sv, qv, qc, qr); as in the reference's set_output, an overwritten output
 * is zero everywhere outside the levels [origin[2] + extent, origin[2] + domain[2] - extent) of the
 * box, i.e. the WHOLE output storage is written. */
int tb200_vertical_advection(int flux_scheme, int staggered_w, const tb200_field *in_w,
                             const tb200_field *in_s, const tb200_field *in_su,
                             const tb200_field *in_sv, tb200_field *out_s, tb200_field *out_su,
                             tb200_field *out_sv, const tb200_field *in_qv,
                             const tb200_field *in_qc, const tb200_field *in_qr,
                             tb200_field *out_qv, tb200_field *out_qc, tb200_field *out_qr,
                             double dz, uint32_t overwrite_flags, const int32_t origin[3],
                             const int32_t domain[3], void *stream);

/* One stage of a tendency stepper around the vertical advection in ONE kernel:
 * out[f] = base[f] + factor * tendency[f](in) on the whole output storage, i.e.
 * tb200_vertical_advection followed by the stage update of
 * src/tasmania/framework/subclasses/tendency_steppers/{forward_euler,rk2,rk3ws}.py
 * (DataArrayDictOperator.fma, src/tasmania/utils/xarrayx.py:L688-L740) without the round trip of
 * the tendencies through memory.  in / base / out = s, su, sv[, qv, qc, qr]; nfields 3 or 6;
 * outputs must alias neither inputs nor base fields (base may alias in). */
int tb200_vertical_advection_step(int flux_scheme, int staggered_w, const tb200_field *in_w,
                                  int nfields, const tb200_field *const *in,
                                  const tb200_field *const *base, tb200_field *const *out,
                                  double dz, double factor, const int32_t origin[3],
                                  const int32_t domain[3], void *stream);


/* ---- implicit (Crank-Nicolson) vertical advection:
 * src/tasmania/isentropic/physics/implicit_vertical_advection.py:L221-L336 with setup_thomas and
 * thomas of src/tasmania/framework/subclasses/subroutine_definitions/cla.py:L42-L108.
 * gamma = dt / (4 dz); staggered_w: in_w on interface levels, averaged onto the main levels.
 * in_q* / out_q* NULL = dry; the water species are advected as s q and returned as mass fractions.
 * dt_tendency = 0: the advected fields (IsentropicImplicitVerticalAdvectionDiagnostic, L221-L336);
 * dt_tendency > 0: the tendencies (x_new - x) / dt_tendency into the same outputs
 * (IsentropicImplicitVerticalAdvectionPrognostic, L793-L919).  2 <= domain[2] <= 256. */
int tb200_implicit_vertical_advection(int staggered_w, const tb200_field *in_w,
                                      const tb200_field *in_s, const tb200_field *in_su,
                                      const tb200_field *in_sv, tb200_field *out_s,
                                      tb200_field *out_su, tb200_field *out_sv,
                                      const tb200_field *in_qv, const tb200_field *in_qc,
                                      const tb200_field *in_qr, tb200_field *out_qv,
                                      tb200_field *out_qc, tb200_field *out_qr, double gamma,
                                      double dt_tendency, const int32_t origin[3],
                                      const int32_t domain[3], void *stream);

/* ---- the global `thomas` stencil: src/tasmania/framework/subclasses/stencil_definitions/
 * cla.py:L33-L62 -- tridiagonal systems a[k] x[k-1] + b[k] x[k] + c[k] x[k+1] = d[k] solved per
 * column over k in [origin[2], origin[2]+domain[2]) with the reference's zero-pivot rule
 * (divide by b[k] where the eliminated diagonal vanishes).  out may alias d (the reference works
 * on a copy of d) but not a, b or c.  At most 256 levels. */
int tb200_thomas(const tb200_field *a, const tb200_field *b, const tb200_field *c,
                 const tb200_field *d, tb200_field *out, const int32_t origin[3],
                 const int32_t domain[3], void *stream);

/* ---- fused dry isentropic stage (the benchmark hot path) ---------------------------
 * One RK stage of IsentropicDynamicalCore.stage_array_call_dry
 * (src/tasmania/isentropic/dynamics/dycore.py:L641-L721) with the relaxed lateral boundary:
 * K1 + irelax(s) + pressure/Exner/Montgomery scans + K2 + irelax(s, su, sv) + Rayleigh
 * damping + velocity diagnosis + outermost layers, in two kernels.  All fields share one
 * storage shape and have unit stride along i.  `scratch_exn`/`scratch_mtg`/`scratch_s` are
 * caller-owned work storages of the same shape.  gamma must be 1 on the nb outermost rings.  hs2d: topography, shape (>=nx, >=ny, 1).
 * rmat1d / gamma2d: damping profile (1,1,nk) and relaxation coefficients (ni,nj,1) given as
 * fields (strides pick the rank).  damp != 0 applies the damping in this stage. */
typedef struct tb200_isentropic_stage {
  int32_t nx, ny, nz, nb; /* numerical grid, boundary layers */
  int32_t flux_scheme;
  int32_t damp;
  double dt;        /* stage time step [s] */
  double dt_full;   /* full time step used by the damping [s] */
  double dx, dy, dz, eps, pt, theta_s;
  double constants[4]; /* pref, rd, g, cp */
  /* Splitting a stage for communication/computation overlap (2-D domain decomposition):
   * part 0 = the whole stage; part 1 = everything except the interior blocks of the momentum
   * kernel, i.e. all results a neighbour's halo needs; part 2 = those interior blocks.
   * rim[4] = number of owned columns / rows next to the (west, east, south, north) edge that
   * must be final after part 1 (0 for an edge without a neighbour). */
  int32_t part;
  int32_t rim[4];
  /* Velocity round trips between the stages of one time step (both default to 0 = the
   * reference's data flow, dycore.py:L702-L721: every stage writes u_new / v_new and the next
   * one reads them as u_int / v_int).
   * derive_uv_in != 0: the advecting velocities are re-diagnosed inside the kernels from s_int,
   *   su_int, sv_int with the formula of velocity_x / velocity_y (dwarfs/diagnostics.py:L219-L272)
   *   instead of being read -- bit-identical whenever u_int / v_int ARE that diagnosis (true for
   *   the output of a previous stage; not for an arbitrary initial state), u_int / v_int are not
   *   touched and may be stale.
   * skip_uv_out != 0: u_new / v_new are not written: an intermediate stage whose consumer sets
   *   derive_uv_in, or the last stage of a step when the caller diagnoses the velocities of the
   *   final state with one tb200_velocity_components pass (what the hosts of this repository do:
   *   cheaper than the in-kernel diagnosis).  With skip_uv_out, scratch_s may be the SAME storage
   *   as s_new: the stage then updates s in place and stores it only where relaxation / damping
   *   changed it. */
  int32_t derive_uv_in;
  int32_t skip_uv_out;
  /* Slow tendencies of s, su, sv (the `s_tnd`, `su_tnd`, `sv_tnd` arguments of the reference's
   * step_forward_euler / step_forward_euler_momentum, prognostics/utils.py:L43-L204, as
   * rk3ws_si.py:L105-L234 passes them): all three or none (NULL).  Same geometry as the other
   * fields; dry stage, part 0 and the default kernel path only. */
  const tb200_field *s_tnd, *su_tnd, *sv_tnd;
  /* Periodic lateral boundary (domain/subclasses/horizontal_boundaries/periodic.py:L98-L122): with
   * periodic != 0, (nx, ny) is the numerical grid of a physical (nx - 2 nb, ny - 2 nb) one, gamma is
   * a field of zeros, and the stage wraps s into its nb ghost layers between the s-step and the
   * column scans -- the `hb.enforce_field(s_new)` of rk3ws_si.py:L184-L189, which with a relaxed
   * boundary is the first relaxation.  Ghost points of su_new / sv_new (and of the water constituents
   * of the moist stage) keep their previous contents: the caller runs the stage with damp = 0 and
   * skip_uv_out and applies `hb.enforce_raw` (tb200_periodic_enforce) and the damping afterwards, in
   * the reference's order (dycore.py:L684-L700).  Part 0, default kernel path, nx, ny >= 4 nb. */
  int32_t periodic;
} tb200_isentropic_stage;

int tb200_isentropic_stage_dry(
    const tb200_isentropic_stage *cfg, const tb200_field *s_now, const tb200_field *su_now,
    const tb200_field *sv_now, const tb200_field *mtg_now, const tb200_field *s_int,
    const tb200_field *su_int, const tb200_field *sv_int, const tb200_field *u_int,
    const tb200_field *v_int, tb200_field *s_new, tb200_field *su_new, tb200_field *sv_new,
    tb200_field *u_new, tb200_field *v_new, const tb200_field *s_ref,
    const tb200_field *su_ref, const tb200_field *sv_ref, const tb200_field *u_ref,
    const tb200_field *v_ref, const tb200_field *gamma, const tb200_field *rmat,
    const tb200_field *hs, tb200_field *scratch_exn, tb200_field *scratch_mtg,
    tb200_field *scratch_s, void *stream);

/* One RK stage of the MOIST dynamical core (src/tasmania/isentropic/dynamics/dycore.py:L723-L843,
 * stage_array_call_moist, without slow tendencies): tb200_isentropic_stage_dry plus the three
 * water constituents, whose reference path is density (dwarfs/diagnostics.py:L400-L416, now and
 * int) -> their share of K1 (prognostics/utils.py:L101-L134) -> mass_fraction (L434-L450, with
 * the stage's s after its first relaxation) -> Relaxed.enforce_raw -- all with clipping.  One
 * extra kernel between the s-step and the column scans does the four for all constituents.
 * q_now / q_int / q_new / q_ref: arrays of three field pointers (mass fractions of water vapour,
 * cloud liquid water, precipitation water) at the start of the step / the stage input / the
 * stage output / the reference state of the lateral boundary.  Same geometry as every other
 * field; default kernel path and part == 0 only. */
int tb200_isentropic_stage_moist(
    const tb200_isentropic_stage *cfg, const tb200_field *s_now, const tb200_field *su_now,
    const tb200_field *sv_now, const tb200_field *mtg_now, const tb200_field *s_int,
    const tb200_field *su_int, const tb200_field *sv_int, const tb200_field *u_int,
    const tb200_field *v_int, tb200_field *s_new, tb200_field *su_new, tb200_field *sv_new,
    tb200_field *u_new, tb200_field *v_new, const tb200_field *s_ref,
    const tb200_field *su_ref, const tb200_field *sv_ref, const tb200_field *u_ref,
    const tb200_field *v_ref, const tb200_field *gamma, const tb200_field *rmat,
    const tb200_field *hs, tb200_field *scratch_exn, tb200_field *scratch_mtg,
    tb200_field *scratch_s, const tb200_field *const *q_now, const tb200_field *const *q_int,
    tb200_field *const *q_new, const tb200_field *const *q_ref, void *stream);

/* Per-kernel timing of the fused stage for the roofline report: with profiling enabled every
 * tb200_isentropic_stage_dry call records CUDA events on its stream around its kernels;
 * tb200_stage_profile_read waits for the last call and returns the durations [ms] of
 * {s-step kernel (A, or S), column-scan kernel (B; 0 when S does both), momentum kernel}. */
int tb200_stage_profile(int enable);
int tb200_stage_profile_read(double ms[3]);
/* 1 if the stage kernels selected by the process environment honour derive_uv_in / skip_uv_out
 * and scratch_s == s_new for columns of nz layers (the default path), 0 if an earlier kernel
 * variant is forced (TB200_S_IMPL, TB200_MV_IMPL, TB200_STAGE_IMPL) or nz > 64: a host then
 * passes 0 for both flags and a separate scratch_s. */
int tb200_stage_lazy_velocities(int nz);

/* ---- halo exchange support (2-D domain decomposition) ------------
 * pack/unpack a box of a field into/from a contiguous buffer (i fastest). */
int tb200_pack_box(const tb200_field *field, double *buffer, const int32_t origin[3],
                   const int32_t domain[3], void *stream);
int tb200_unpack_box(tb200_field *field, const double *buffer, const int32_t origin[3],
                     const int32_t domain[3], void *stream);
/* the same box of `nfields` (<= TB200_HALO_MAX_FIELDS) fields in ONE launch; message layout
 * [field][k][j][i], i fastest.  One side of a halo exchange = one pack + one unpack. */
#define TB200_HALO_MAX_FIELDS 8
int tb200_halo_pack(const tb200_field *const *fields, int nfields, double *buffer,
                    const int32_t origin[3], const int32_t domain[3], void *stream);
int tb200_halo_unpack(const tb200_field *const *fields, int nfields, const double *buffer,
                      const int32_t origin[3], const int32_t domain[3], void *stream);

/* ---- peer-to-peer halo transport over NVLink / NVSwitch (one process per GPU).  No reference
 * counterpart (the reference is single-process; north_star: "NCCL/P2P halo exchange over NVLink").
 * A rank allocates its receive buffers and arrival counters with tb200_p2p_alloc, exports the
 * allocation (CUDA IPC handle, 64 bytes, passed to the neighbours by any host channel) and
 * imports the neighbours' ones.  tb200_halo_push packs the send slabs of up to
 * TB200_HALO_MAX_SIDES sides STRAIGHT into the neighbours' receive buffers (peer stores) and raises
 * their arrival counters; tb200_halo_pull waits for the own counters and unpacks.  Receive buffers hold two
 * slots of slot_doubles (exchange q uses slot q & 1; see csrc/halo.cu for why two suffice);
 * `channel` = TB200_P2P_CHANNEL_BYTES of zeroed device memory per phase (sequence numbers, kept
 * on the device so that the launches can be captured in a CUDA graph).  Both sides of a pair must
 * issue their pushes / pulls in the same order. */
#define TB200_HALO_MAX_SIDES 8 /* four faces + four corner blocks in one launch */
#define TB200_P2P_HANDLE_BYTES 64
#define TB200_P2P_CHANNEL_BYTES 32
typedef struct {
  double *remote_buffer;    /* neighbour's receive buffer for the slab sent to it (peer-mapped) */
  uint64_t *remote_counter; /* neighbour's arrival counter of that buffer (peer-mapped) */
  double *local_buffer;     /* own receive buffer of this side */
  uint64_t *local_counter;  /* own arrival counter, raised by the neighbour */
  int64_t slot_doubles;     /* capacity of one slot */
  int32_t send_origin[2], recv_origin[2], extent[2]; /* (i, j) of the outgoing / incoming slab */
} tb200_halo_side;
int tb200_halo_push(const tb200_field *const *fields, int nfields, const tb200_halo_side *sides,
                    int nsides, void *channel, int k0, int nk, void *stream);
int tb200_halo_pull(const tb200_field *const *fields, int nfields, const tb200_halo_side *sides,
                    int nsides, void *channel, int k0, int nk, void *stream);
int tb200_p2p_alloc(size_t bytes, void **ptr);   /* zero-filled, exportable device memory */
int tb200_p2p_free(void *ptr);
int tb200_p2p_export(void *ptr, unsigned char handle[TB200_P2P_HANDLE_BYTES]);
int tb200_p2p_import(const unsigned char handle[TB200_P2P_HANDLE_BYTES], void **ptr);
int tb200_p2p_release(void *ptr);                /* undo tb200_p2p_import */
/* *error != 0: a pull gave up waiting for its peer (~20 s); the results are invalid */
int tb200_p2p_channel_error(const void *channel, int *error);

/* ---- Coriolis / Smagorinsky fused with the stage update of a tendency stepper (b200 only; the
 * reference runs the component's stencil into tendency storages and then the `fma` stencil over the
 * whole storages: src/tasmania/framework/subclasses/tendency_steppers/{forward_euler,rk2,rk3ws}.py
 * over src/tasmania/utils/xarrayx.py:L688-L740):
 *   out = base + factor * tendency on [origin, origin + domain),  out = base + factor * 0 on the rest
 * of the `full` box [0, full) (the storages' shape).  Tendency formulas and argument meaning as in
 * tb200_coriolis / tb200_smagorinsky (in_s NULL = the non-isentropic Smagorinsky2d). */
int tb200_coriolis_step(const tb200_field *in_su, const tb200_field *in_sv,
                        const tb200_field *base_su, const tb200_field *base_sv, tb200_field *out_su,
                        tb200_field *out_sv, double f, double factor, const int32_t origin[3],
                        const int32_t domain[3], const int32_t full[3], void *stream);
int tb200_smagorinsky_step(const tb200_field *in_s, const tb200_field *in_a, const tb200_field *in_b,
                           const tb200_field *base_a, const tb200_field *base_b, tb200_field *out_a,
                           tb200_field *out_b, double dx, double dy, double cs, double factor,
                           const int32_t origin[3], const int32_t domain[3], const int32_t full[3],
                           void *stream);

/* ---- self-test of the kernels' own correctly rounded division (csrc/common.cuh: qdiv) against
 * the compiler's IEEE division on `count` generated operand pairs (random, guard-crossing and
 * hard-case classes); *mismatches (device memory, zeroed by the caller) receives the number of
 * pairs whose results differ.  Test infrastructure for tests/test_gpu_division.py. */
int tb200_selftest_division(uint64_t count, uint64_t seed, uint64_t *mismatches, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TASMANIA_B200_H */
