/*
 * ssb200.h - C ABI of libssb200.so, the B200-native (sm_100a) hot path of streamsculptor.
 *
 * The reference (jnibauer/streamsculptor) has NO native boundary: its hot path is Python/JAX that funnels into
 * diffrax.diffeqsolve inside one jitted XLA program.  The entry points below are therefore what
 * an XLA-FFI custom call (or any ctypes / cgo / JNI binding) for that path binds to; each one cites the reference
 * interface it replaces (paths relative to /root/reference/streamsculptor/).  INTEGRATION.md shows the binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C, no torch / JAX types.  All array arguments of the *_f64 entry points are DEVICE pointers owned by the
 *     caller; work is enqueued on `stream` (a cudaStream_t passed as void*) and the call returns without synchronising.
 *     The *_host entry points take HOST pointers, do the H2D/D2H copies themselves and synchronise before returning.
 *   - no global mutable state: entry points are re-entrant; potential descriptions are passed by value per call.
 *   - return value: 0 = enqueued, <0 = ssb_status error (bad argument, unsupported potential, CUDA error).
 *     NUMERICAL failure is per orbit in status[]: 0 ok, 1 max_steps reached, 2 non-finite; rows never saved are +inf
 *     (diffrax throw=False semantics, main.py:136).
 *   - units: kpc, Myr, Msun; all real data IEEE fp64.  Arrays are C-contiguous in the layout the reference's Python
 *     API uses ([N,6] phase-space rows, [N,M,6] saved trajectories, [N,Nsh,12] responses).
 */
#ifndef SSB200_H
#define SSB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_ABI_VERSION 2
#define SSB_MAX_COMP 12
#define SSB_MAX_TRACK 4
#define SSB_MAX_SUBHALO_SETS 2
#define SSB_MAX_PSETS 1

typedef enum {
    SSB_OK = 0,
    SSB_ERR_ARG = -1,         /* null pointer / negative size / inconsistent shapes */
    SSB_ERR_UNSUPPORTED = -2, /* component type, profile or solver this library does not implement */
    SSB_ERR_CUDA = -3,        /* CUDA runtime error (ssb_last_error() has the text) */
    SSB_ERR_SCRATCH = -4      /* scratch buffer too small (see ssb_scratch_bytes) */
} ssb_status;

/* ---- potential program: a flat sum of components (Potential_Combine, potential.py:1279-1296) ------------------- */
typedef enum {
    SSB_NFW = 0,          /* potential.py:74-84    p = {G*m, r_s}                      */
    SSB_HERNQUIST = 1,    /* potential.py:132-138  p = {G*m, r_s, soft}                */
    SSB_MIYAMOTO = 2,     /* potential.py:66-72    p = {G*m, a, b}                     */
    SSB_PLUMMER = 3,      /* potential.py:124-130  p = {G*m, r_s}                      */
    SSB_ISOCHRONE = 4,    /* potential.py:114-122  p = {G*m, a}                        */
    SSB_TRIAXNFW = 5,     /* potential.py:86-97    p = {G*m, r_s, q1, q2, q3}          */
    SSB_UNIFORM_ACC = 6,  /* potential.py:480-502  gradient = d(velocity track)/dt ; track = velocity table */
    SSB_SUBHALOS = 7,     /* potential.py:802-850, 1161-1213  sh = index into ssb_potential.sh              */
    SSB_PERTURBERS = 8,   /* a SET of moving spheres on tabulated tracks (sum of TimeDepTranslatingPotential components,
                             potential.py:448-462, as the restricted N-body / LMC set-ups build them - BASELINE config 5:
                             100 live perturbers); sh = index into ssb_potential.pset                        */
    SSB_BAR = 9,          /* potential.py:178-198  Long & Murali bar rotating with Omega: p = {G*m, a, b, c, Omega}       */
    SSB_DEHNEN_BAR = 10   /* potential.py:200-222  p = {alpha, v0, R0, Rb, phib, Omega}                                    */
} ssb_comp_type;

typedef struct {
    int32_t type;   /* ssb_comp_type */
    int32_t track;  /* >= 0: evaluate at x - c(t), c = tracks[track] (TimeDepTranslatingPotential, potential.py:448-462); -1: static */
    int32_t sh;     /* SSB_SUBHALOS: which subhalo set */
    int32_t growth; /* GrowingPotential (potential.py:464-477): 0 = none, g > 0: the component's mass scale is multiplied by the FIRST column of
                       tracks[g - 1] evaluated at t (a tabulated growth factor).  Not for SSB_UNIFORM_ACC / SSB_SUBHALOS components. */
    double p[8];
} ssb_component;

typedef enum {
    SSB_TRACK_LINEAR = 0, /* RegularGridInterpolator(method='linear', fill_value=None): potential.py:581-600 (linear extrapolation) */
    SSB_TRACK_CUBIC = 1   /* interpax.Interpolator1D(method='cubic'): streamhelpers.py:520, perturbative.py:642 (NaN outside knots)  */
} ssb_track_kind;

typedef struct {
    int32_t kind;     /* ssb_track_kind */
    int32_t n;        /* knots */
    const double* t;  /* [n] increasing (device) */
    const double* y;  /* [n,3] (device) */
    const double* s;  /* [n,3] knot slopes for SSB_TRACK_CUBIC (device; fill with ssb_track_slopes_f64), NULL for linear */
    double t_first;   /* reserved, set to 0: the kernels fill t[0] ... */
    double inv_dt;    /* ... and (n-1)/(t[n-1]-t[0]) in their shared-memory copy (segment guess without a division per force) */
} ssb_track;

typedef enum { SSB_PROFILE_PLUMMER = 0, SSB_PROFILE_HERNQUIST = 1, SSB_PROFILE_NFW = 2 } ssb_profile;

/* N_sh subhalos on straight lines x0 + v (t - t0), active iff |t - t0| < t_window (strict; potential.py:826). */
typedef struct {
    int32_t n;
    int32_t profile;    /* ssb_profile: SubhaloLinePotential = Plummer; ..Custom_fromFunc = any func(m, r_s) - Hernquist in production */
    double G;
    const double* m;    /* [n]   (device) */
    const double* rs;   /* [n]   */
    const double* x0;   /* [n,3] */
    const double* v;    /* [n,3] */
    const double* t0;   /* [n]   */
    const double* tw;   /* [n]   t_window per subhalo (broadcast a scalar on the host side) */
} ssb_subhalos;

/* n moving spheres of one profile whose centres share ONE time grid (linear interpolation, linear extrapolation outside, as
 * SSB_TRACK_LINEAR): one packed table instead of n components + n tracks, so that programs are not limited to SSB_MAX_COMP /
 * SSB_MAX_TRACK moving bodies, and kernels whose particles share the stage times (ssb_shared_step_orbits_f64) interpolate every
 * centre once per stage and CTA. */
typedef struct {
    int32_t n;          /* perturbers */
    int32_t n_knots;    /* knots of the shared time grid (>= 2) */
    int32_t profile;    /* ssb_profile of every perturber */
    int32_t _pad;
    const double* t;    /* [n_knots] increasing (device) */
    const double* y;    /* [n_knots, n, 3] centres */
    const double* GM;   /* [n] G m */
    const double* rs;   /* [n] scale radii */
} ssb_perturbers;

typedef struct {
    int32_t n_comp, n_track, n_sh, n_pset;
    ssb_component comp[SSB_MAX_COMP];
    ssb_track track[SSB_MAX_TRACK];
    ssb_subhalos sh[SSB_MAX_SUBHALO_SETS];
    ssb_perturbers pset[SSB_MAX_PSETS];
} ssb_potential;

/* ---- solver control: diffrax.PIDController(rtol, atol, dtmin, dtmax, force_dtmin=True) + max_steps (main.py:144-162) */
typedef struct {
    int32_t solver;     /* 5 = Dopri5, 8 = Dopri8 */
    int32_t max_steps;  /* accepted + rejected */
    double rtol, atol;
    double dtmin;       /* steps at dt <= dtmin are always accepted */
    double dtmax;       /* +inf = None */
} ssb_ctrl;

int ssb_abi_version(void);
const char* ssb_last_error(void); /* thread-local text of the last SSB_ERR_CUDA / SSB_ERR_ARG */
unsigned long long ssb_launch_count(void); /* kernels this library has launched in this process so far (bench.py: gpu_launches) */

/* A1/A5  Potential.potential / gradient / jacobian_force (main.py:37-65) at n points.
 * Any of phi[n], grad[n,3], hess[n,3,3] may be NULL.  SSB_UNIFORM_ACC contributes to grad only. */
int ssb_potential_eval_f64(const ssb_potential* pot, int64_t n, const double* xyz, const double* t,
                           double* phi, double* grad, double* hess, void* stream);

/* per-subhalo values of one subhalo set at ONE point: potential_per_SH and its jacfwd (potential.py:832-850;
 * perturbative.py:40-41, 695-696).  dradius != 0 evaluates d/dr_s of the profile (potential.py:852-904, 1215-1268).
 * phi[n_sh], grad[n_sh,3]. */
int ssb_subhalo_eval_f64(const ssb_subhalos* sh, int dradius, const double* xyz /*[3] host*/, double t,
                         double* phi, double* grad, void* stream);

/* knot slopes of a cubic track (interpax 'cubic' derivative rule).  t[n], y[n,3] -> s[n,3]. */
int ssb_track_slopes_f64(int64_t n, const double* t, const double* y, double* s, void* stream);
/* evaluate a track and its time derivative at nq times: out[nq,3], dout[nq,3] (dout may be NULL) */
int ssb_track_eval_f64(const ssb_track* tr, int64_t nq, const double* tq, double* out, double* dout, void* stream);

/* A3/A4  Potential.integrate_orbit / integrate_orbit_batch_vmapped (main.py:125-202): N independent adaptive solves.
 *   w0[N,6]; t0[N], t1[N] (t1 < t0 integrates backwards); ts: save times, [N,M] if ts_per_orbit else [M] shared,
 *   monotone from t0 towards t1; ys[N,M,6]; status[N]; nsteps[N,3] = {attempted, accepted, rejected}. */
int ssb_orbit_integrate_f64(const ssb_potential* pot, int64_t N, const double* w0, const double* t0, const double* t1,
                            const double* ts, int32_t M, int32_t ts_per_orbit, ssb_ctrl ctrl,
                            double* ys, int32_t* status, int32_t* nsteps, void* stream);

/* A7 (serial part)  ONE orbit saved at M times with dense output (main.py:289: the progenitor at every stripping time).
 * Steps are taken by one thread and recorded; the M interpolations run in parallel.  scratch >= ssb_scratch_bytes(). */
int ssb_orbit_dense_f64(const ssb_potential* pot, const double* w0 /*[6] device*/, double t0, double t1,
                        const double* ts, int64_t M, ssb_ctrl ctrl, double* ys /*[M,6]*/, int32_t* status /*[1]*/,
                        int32_t* nsteps /*[3]*/, void* scratch, size_t scratch_bytes, void* stream);
size_t ssb_scratch_bytes(int32_t max_steps);
/* A3/A8  batched dense solutions: integrate_orbit(dense=True) under vmap (main.py:139-162, dense branch) as used by
 * gen_stream_scan_dense / gen_stream_vmapped_dense (main.py:376-430) and streamhelpers.eval_dense_stream (streamhelpers.py:23-53).
 * Every accepted step of every orbit is recorded (rec_cap slots of 512 B per orbit; more accepted steps than slots -> status 1);
 * ssb_orbit_record_eval_f64 then evaluates all N interpolants at one common time (per_orbit = 0) or at tq[i]; +inf outside the
 * integrated interval.  recs >= ssb_record_bytes(N, rec_cap) bytes. */
int ssb_orbit_record_f64(const ssb_potential* pot, int64_t N, const double* w0, const double* t0, const double* t1, ssb_ctrl ctrl, int32_t rec_cap,
                         void* recs, size_t rec_bytes, int32_t* status, int32_t* nsteps, void* stream);
int ssb_orbit_record_eval_f64(int32_t solver, int64_t N, const void* recs, int32_t rec_cap, const double* tq, int32_t per_orbit, double* ys, void* stream);
size_t ssb_record_bytes(int64_t N, int32_t rec_cap);
/* Diagnostics for the lock-step parity tests (tests/test_gpu_lockstep.py): the controller's view of every step ATTEMPT of N adaptive
 * solves with the same stepper code as ssb_orbit_integrate_f64 - trace[N,trace_cap,4] = {t_prev, dt, err, keep} in mirrored time
 * (t * direction; rows beyond an orbit's attempt count are left untouched) - and the final states yfin[N,6].  What it mirrors in the
 * reference: the (tprev, tnext, y_error norm, keep_step) carried by diffrax's adaptive loop under main.py:139-162. */
int ssb_orbit_trace_f64(const ssb_potential* pot, int64_t N, const double* w0, const double* t0, const double* t1, ssb_ctrl ctrl,
                        int32_t trace_cap, double* trace, double* yfin, int32_t* status, int32_t* nsteps, void* stream);
/* Solution.evaluate(t) of a dense=True solve (main.py:131, 141): interpolate M more times inside the steps recorded in `scratch` by a
 * previous ssb_orbit_dense_f64 call with the same solver; ys[M,6] (+inf outside the integrated interval). */
int ssb_orbit_dense_eval_f64(int32_t solver, const void* scratch, const double* ts, int64_t M, double* ys, void* stream);

/* A6  Potential.release_model vmapped over stripping times (main.py:209-306).
 *   prog[N,6] progenitor phase-space at t[N]; Msat[N]; idx[N] the stripping index i that seeds jax.random
 *   (main.py:223-228); kvals[8] (host) = {kr, kvphi, kz, kvz, sigma_kr, sigma_kvphi, sigma_kz, sigma_kvz};
 *   normals[N,4] optional externally supplied standard normals (NULL = reproduce the jax threefry recipe);
 *   outputs pos_lead, pos_trail, vel_lead, vel_trail each [N,3]. */
int ssb_release_spray_f64(const ssb_potential* pot, double G, int64_t N, const double* prog, const double* Msat,
                          const int64_t* idx, const double* t, int64_t seed, const double* kvals, const double* normals,
                          double* pos_lead, double* pos_trail, double* vel_lead, double* vel_trail, void* stream);

/* A9  release_model_Chen25 vmapped over stripping times (streamhelpers.py:352-459): per release 6 correlated normals
 * (jax.random.multivariate_normal(key_i, mean, cov, method='svd'), key_i = jax.random.split(key, N)[i]) turned into offsets
 * in units of the tidal radius / escape velocity.  key[2] (host) = the jax PRNG key words; mean[6], factor[36] (host, row
 * major) with sample = mean + factor @ z; normals[N,6] optional externally supplied standard normals (then key may be NULL). */
int ssb_release_chen25_f64(const ssb_potential* pot, double G, int64_t N, const double* prog, const double* Msat, const double* t,
                           const uint32_t* key, const double* mean, const double* factor, const double* normals,
                           double* pos_lead, double* pos_trail, double* vel_lead, double* vel_trail, void* stream);

/* A15  jacfwd(release_model) (BaseStreamModel.release_func_jacobian, perturbative.py:281-296): d(pos, vel of the lead / trail
 * particle) / d(progenitor x, v) at every stripping time; needs the third derivatives of Phi (closed forms on the device).
 * Arguments as ssb_release_spray_f64; jac[N,2,6,6]. */
int ssb_release_jacobian_f64(const ssb_potential* pot, double G, int64_t N, const double* prog, const double* Msat,
                             const int64_t* idx, const double* t, int64_t seed, const double* kvals, const double* normals,
                             double* jac, void* stream);
/* third derivatives d^3 Phi / dx_i dx_j dx_k at n points: third[n,3,3,3] (nested jacfwd of the gradient, fields.py:278-283) */
int ssb_potential_third_f64(const ssb_potential* pot, int64_t n, const double* xyz, const double* t, double* third, void* stream);

/* A8  Potential.gen_stream_vmapped (main.py:343-368) as one enqueue: progenitor orbit at ts[Nts] (dense), release at
 * every ts[i], then 2 (Nts-1) independent solves from ts[i] to ts[Nts-1]; lead[Nts-1,6], trail[Nts-1,6].
 * pot_release: potential used by release_model (== pot for gen_stream_vmapped; the base potential for
 * gen_stream_vmapped_with_pert, streamhelpers.py:56-116).  Only the n_local particles i = i_begin + k*i_stride
 * (k = 0..n_local-1, all < Nts-1) are integrated - interleaved multi-GPU sharding, every rank sees the same mix of
 * integration spans; outputs are indexed by k.  scratch >= ssb_stream_scratch_bytes().
 * Streams with n_local >= 32768 whose lead and trail are the halves of one [2, n_local, 6] buffer run as up to four concurrent parts
 * (the earliest particles behind a cut-short progenitor solve on `stream`, the later ones behind longer solves on library-owned
 * streams that fork from and join `stream` through events): same results, the serial progenitor solve leaves the critical path.
 * Work enqueued on `stream` after the call is ordered after all parts. */
int ssb_gen_stream_f64(const ssb_potential* pot, const ssb_potential* pot_release, double G, int64_t Nts,
                       const double* ts, const double* prog_w0 /*[6]*/, const double* Msat /*[Nts]*/, int64_t seed,
                       const double* kvals /*host[8]*/, const double* normals /*[Nts,4] or NULL*/, ssb_ctrl ctrl,
                       int64_t i_begin, int64_t i_stride, int64_t n_local, double* lead, double* trail,
                       int32_t* status /*[2,n_local]*/, int32_t* nsteps /*[2,n_local,3]*/, void* scratch, size_t scratch_bytes,
                       void* stream);
size_t ssb_stream_scratch_bytes(int64_t Nts, int32_t max_steps);

/* A12-A14  compute_perturbation_OTF (perturbative.py:101-135, 425-454, 726-755) with the field
 * MassRadiusPerturbation_OTF (fields.py:159-206): per particle ONE coupled ODE [w(6), D(n_sh,12)] whose step-size
 * controller sees the RMS error over all 6 + 12 n_sh components; final state kept.
 *   w0[N,6]; D0[N,n_sh,12] or NULL (= zeros, perturbative.py:712); t0[N]; t1 common end time;
 *   wout[N,6]; Dout[N,n_sh,12]; scratch >= ssb_response_scratch_bytes(n_sh). */
int ssb_linear_response_f64(const ssb_potential* pot_base, const ssb_subhalos* sh, int64_t N, const double* w0,
                            const double* D0, const double* t0, double t1, ssb_ctrl ctrl, double* wout, double* Dout,
                            int32_t* status, int32_t* nsteps, void* scratch, size_t scratch_bytes, void* stream);
size_t ssb_response_scratch_bytes(int32_t n_sh);
/* Same field for ONE trajectory with SaveAt(ts): integrate_field(w0=[w, D], ts, backwards_int=...) as used for the backward
 * progenitor response (perturbative.py:53-60, 394-398, 704-711).  w0[6], D0[n_sh,12] or NULL, t0[1] (device), ts[M] monotone
 * from t0 towards t1; ws[M,6], Ds[M,n_sh,12]; scratch >= ssb_response_saveat_scratch_bytes(n_sh). */
int ssb_linear_response_saveat_f64(const ssb_potential* pot_base, const ssb_subhalos* sh, const double* w0, const double* D0,
                                   const double* t0, double t1, const double* ts, int32_t M, ssb_ctrl ctrl, double* ws, double* Ds,
                                   int32_t* status, int32_t* nsteps, void* scratch, size_t scratch_bytes, void* stream);
size_t ssb_response_saveat_scratch_bytes(int32_t n_sh);
/* A16  second-order mass response: fields.MassRadiusPerturbation_OTF_SecondOrder (fields.py:260-320) integrated per particle as in
 * GenerateMassRadiusPerturbation_Chen25.compute_perturbation_second_order_OTF (perturbative.py:757-772): state
 * [w(6), D(n_sh,12), E(n_sh,6)], one controller over all 6 + 18 n_sh components, final state kept.
 * D0 / E0 may be NULL (= zeros, perturbative.py:715).  scratch >= ssb_second_order_scratch_bytes(n_sh). */
int ssb_second_order_response_f64(const ssb_potential* pot_base, const ssb_subhalos* sh, int64_t N, const double* w0, const double* D0,
                                  const double* E0, const double* t0, double t1, ssb_ctrl ctrl, double* wout, double* Dout, double* Eout,
                                  int32_t* status, int32_t* nsteps, void* scratch, size_t scratch_bytes, void* stream);
/* The same solve with PER-PARTICLE end times t1[N] (t1[i] < t0[i] integrates backwards): N copies of the progenitor, each integrated from
 * the observed position back to one stripping time, give the backward-integrated progenitor field at every stripping time that
 * GenerateMassRadiusPerturbation_CustomBase_SecondOrder maps into its perturbation ICs (perturbative.py:519-552). */
int ssb_second_order_response_ends_f64(const ssb_potential* pot_base, const ssb_subhalos* sh, int64_t N, const double* w0, const double* D0,
                                       const double* E0, const double* t0, const double* t1, ssb_ctrl ctrl, double* wout, double* Dout,
                                       double* Eout, int32_t* status, int32_t* nsteps, void* scratch, size_t scratch_bytes, void* stream);
size_t ssb_second_order_scratch_bytes(int32_t n_sh);
/* RHS of the second-order field at one state y = [w(6), D(n_sh,12), E(n_sh,6)] (fields.py:289-320), for unit tests */
int ssb_second_order_term_f64(const ssb_potential* pot_base, const ssb_subhalos* sh, double t, const double* y, double* dy, void* stream);
/* RHS of that field at one state (fields.py:175-206): y[6+12 n_sh] -> dy (device pointers), for unit tests */
int ssb_response_term_f64(const ssb_potential* pot_base, const ssb_subhalos* sh, double t, const double* y, double* dy,
                          void* stream);

/* A17  N tracers integrated as ONE ODE state with a shared step-size controller: RestrictedNbody_generator.term
 * (RestrictedNbody.py:93-106) integrated by fields.integrate_field (RestrictedNbody.py:131, fields.py:85-98).  `pot` = external
 * potential + progenitor monopole on its track (a translating component).  The RMS error norm runs over all 6N components;
 * w0/wout [N,6]; status[1]; nsteps[3] = steps, accepted, rejected.  scratch >= ssb_shared_scratch_bytes(N).  The call polls a
 * device flag between batches of step attempts, so it synchronises the stream. */
int ssb_shared_step_orbits_f64(const ssb_potential* pot, int64_t N, const double* w0, double t0, double t1, ssb_ctrl ctrl, double* wout,
                               int32_t* status, int32_t* nsteps, void* scratch, size_t scratch_bytes, void* stream);
size_t ssb_shared_scratch_bytes(int64_t N);
/* A17  fields.Nbody_field (fields.py:115-155) through integrate_field: N <= 1024 live bodies, softened all-pairs gravity
 * (softening eps) + optional external potential (ext NULL or n_comp == 0 -> none), ONE ODE state (N,6), SaveAt(ts[M]) ->
 * ys[M,N,6].  scratch >= ssb_nbody_scratch_bytes(N). */
int ssb_nbody_integrate_f64(const ssb_potential* ext, int32_t N, const double* masses, double G, double eps, const double* w0, double t0, double t1,
                            const double* ts, int32_t M, ssb_ctrl ctrl, double* ys, int32_t* status, int32_t* nsteps, void* scratch,
                            size_t scratch_bytes, void* stream);
size_t ssb_nbody_scratch_bytes(int32_t N);
/* Nbody_field.term at one state (fields.py:134-155): y[N,6] -> dy[N,6]; scratch >= 3 N doubles */
int ssb_nbody_term_f64(const ssb_potential* ext, int32_t N, const double* masses, double G, double eps, double t, const double* y, double* dy,
                       void* scratch, size_t scratch_bytes, void* stream);

/* A16 / N3  variational (tangent) equations along unperturbed orbits: examples/higher_order_variationalEqn.ipynb cell 3
 * (second_order_field.term through fields.CustomField, fields.py:362-377; integrate_field, fields.py:35-99).  Per particle ONE
 * solve of [w(6), M(6,6) = dw/dw_init, M2(6,6,6) = d2w/dw_init^2] (order 1: w and M only) from t0[i] to t1 with one controller over
 * all 42 (258) components; final state kept.  M0 NULL = identity, M20 NULL = zeros.  M is the state-transition matrix, i.e. the
 * forward-mode Jacobian of integrate_orbit (main.py:149-162 with adjoint=ForwardMode()) with respect to w0.
 * Layouts: M[N,6,6] row-major M[a][k] = dw_a/dw0_k; M2[N,6,6,6] M2[a][k][l].  status[N], nsteps[N,3]. */
int ssb_variational_f64(const ssb_potential* pot, int32_t order, int64_t N, const double* w0, const double* M0, const double* M20, const double* t0,
                        double t1, ssb_ctrl ctrl, double* wout, double* Mout, double* M2out, int32_t* status, int32_t* nsteps, void* stream);
/* the variational field at one state y[42 | 258] -> dy (device pointers), for unit tests */
int ssb_variational_term_f64(const ssb_potential* pot, int32_t order, double t, const double* y, double* dy, void* stream);

/* ---- host-pointer conveniences (H2D, launch, D2H, synchronise) - what a CPU-side plugin call looks like ----------
 * Zero-copy outputs: if the final-state outputs are PINNED (page-locked, mapped) host memory the orbit kernel writes them
 * directly over PCIe while the remaining orbits integrate, and no D2H copy follows:
 *   ssb_gen_stream_host      - lead and trail are the two halves of ONE pinned [2, n_local, 6] buffer (trail == lead + 6 n_local),
 *                              status [2 n_local] and nsteps [2 n_local, 3] pinned too;
 *   ssb_orbit_integrate_host - final-state mode (M == 1, ts_per_orbit, ts == t1) with ys, status, nsteps pinned.
 * Any other combination (pageable memory, separate lead / trail allocations) takes the staged path: results in HBM, then
 * cudaMemcpyAsync.  Both paths give identical bytes.  SSB_HOST_ZEROCOPY=0 in the environment forces the staged path. */
int ssb_orbit_integrate_host(const ssb_potential* pot_hostptrs, int64_t N, const double* w0, const double* t0,
                             const double* t1, const double* ts, int32_t M, int32_t ts_per_orbit, ssb_ctrl ctrl,
                             double* ys, int32_t* status, int32_t* nsteps);
int ssb_gen_stream_host(const ssb_potential* pot_hostptrs, const ssb_potential* pot_release_hostptrs, double G, int64_t Nts,
                        const double* ts, const double* prog_w0, const double* Msat, int64_t seed, const double* kvals,
                        const double* normals, ssb_ctrl ctrl, int64_t i_begin, int64_t i_stride, int64_t n_local, double* lead,
                        double* trail, int32_t* status, int32_t* nsteps);
int ssb_linear_response_host(const ssb_potential* pot_base_hostptrs, const ssb_subhalos* sh_hostptrs, int64_t N,
                             const double* w0, const double* D0, const double* t0, double t1, ssb_ctrl ctrl,
                             double* wout, double* Dout, int32_t* status, int32_t* nsteps);

/* measured-peak helper for the roofline denominator: runs a dependent-chain-free DFMA loop on every SM and returns
 * the achieved fp64 FLOP/s (2 flops per DFMA); used by bench.py, never by the product path. */
int ssb_fp64_peak_probe(int iters, double* flops_per_s, void* stream);

#ifdef __cplusplus
}
#endif
#endif
