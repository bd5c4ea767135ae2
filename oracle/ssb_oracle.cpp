// TEST INFRASTRUCTURE ONLY - CPU oracle for the streamsculptor hot path.
//
// This library is the CHECKER for the CUDA product path (streamsculptor_b200/csrc).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it; nothing in the
// product package imports, links or executes it.  It restates, on the CPU, the algorithm the reference
// executes (citations: /root/reference/streamsculptor/<file>:<line>):
//   potentials + AD derivatives   orc_potential.h / orc_ad.h   (potential.py, main.py:37-112)
//   diffrax driver                orc_solver.h                 (main.py:139-162, fields.py:85-98)  [3P]
//   jax.random (threefry)         below                        (main.py:220-228, 263-266)          [3P]
//   release_model / gen_stream    below                        (main.py:209-368)
//   linear-response field         below                        (fields.py:159-206, perturbative.py:101-135,726-755)
// PARITY STATUS: jax/diffrax cannot run in this environment, so the oracle is pinned only by the
// reference's notebook goldens (tests/test_oracle_goldens.py): G (G1), the force (D1), the solver / controller / dense output
// (D2-D5), and - through the printed release Jacobian D8 - release_model, its derivative and the jax.random recipe, to the
// 9 printed digits; OC, the reference's printed Orphan-Chenab stream, pins the whole stream pipeline end to end (36 numbers to
// 4e-7, nothing fitted); B1, the printed final velocities of 1000 adaptive Dopri8 orbits, to 5e-10.  Pinned to the printed digits;
// beyond them (1e-10 against a live diffrax run) there is nothing to compare with.
#include <cstdint>
#include <cstring>
#include <vector>
#include <atomic>
#include <thread>

#include "orc_potential.h"
#include "orc_solver.h"

using namespace orc;

// dynamic-schedule parallel loop over independent particles (std::thread; no OpenMP dependency)
template <class Fn> static void parallel_for(int N, int nthreads, int chunk, Fn fn) {
    if (nthreads <= 1 || N <= chunk) { for (int i = 0; i < N; ++i) fn(i); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int w = 0; w < nthreads; ++w)
        pool.emplace_back([&]() {
            for (;;) { int b = next.fetch_add(chunk); if (b >= N) break; int e = std::min(N, b + chunk); for (int i = b; i < e; ++i) fn(i); }
        });
    for (auto& th : pool) th.join();
}

// =============================================================================================
// jax.random, threefry2x32 (jax 0.4.38, jax_threefry_partitionable=False, x64 enabled)   [3P]
// =============================================================================================
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static void threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t* o0, uint32_t* o1) {
    static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
    uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
    uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
    for (int g = 0; g < 5; ++g) {
        for (int r = 0; r < 4; ++r) { x0 += x1; x1 = rotl32(x1, R[g & 1][r]); x1 ^= x0; }
        x0 += ks[(g + 1) % 3];
        x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
    }
    *o0 = x0; *o1 = x1;
}
struct Key { uint32_t a, b; };
static Key prng_key(int64_t seed) {                                    // jax.random.PRNGKey(int64)
    uint64_t u = (uint64_t)seed;
    return Key{(uint32_t)(u >> 32), (uint32_t)(u & 0xFFFFFFFFu)};
}
static void key_split2(Key k, Key* k1, Key* k2) {                      // jax.random.split(key) -> (2,2)
    uint32_t a0, b0, a1, b1;
    threefry2x32(k.a, k.b, 0, 2, &a0, &b0);
    threefry2x32(k.a, k.b, 1, 3, &a1, &b1);
    *k1 = Key{a0, a1}; *k2 = Key{b0, b1};
}
static void random_bits64(Key k, int n, uint64_t* out) {               // _threefry_random_bits(key, 64, (n,))
    for (int j = 0; j < n; ++j) {
        uint32_t o0, o1;
        threefry2x32(k.a, k.b, (uint32_t)j, (uint32_t)(n + j), &o0, &o1);
        out[j] = ((uint64_t)o0 << 32) | (uint64_t)o1;
    }
}
static void randint5(Key key, int64_t minval, int64_t maxval, int64_t* out) {   // jax.random.randint(key,(5,),minval,maxval) int64
    Key k1, k2; key_split2(key, &k1, &k2);
    uint64_t hi[5], lo[5];
    random_bits64(k1, 5, hi); random_bits64(k2, 5, lo);
    uint64_t span = (uint64_t)(maxval - minval);
    if (maxval <= minval) span = 1;
    uint64_t mult = ((uint64_t)1 << 32) % span;
    mult = (mult * mult) % span;
    for (int j = 0; j < 5; ++j) {
        uint64_t off = ((hi[j] % span) * mult + (lo[j] % span)) % span;
        out[j] = minval + (int64_t)off;
    }
}
static double erfinv_f64(double x) {
    // Giles' single-precision polynomial as the seed, then Halley steps on erf(y) - x (converges to ~1 ulp).
    if (x <= -1.0) return -std::numeric_limits<double>::infinity();
    if (x >= 1.0) return std::numeric_limits<double>::infinity();
    double w = -std::log((1.0 - x) * (1.0 + x)), p;
    if (w < 5.0) {
        w -= 2.5;
        p = 2.81022636e-08; p = 3.43273939e-07 + p * w; p = -3.5233877e-06 + p * w; p = -4.39150654e-06 + p * w;
        p = 0.00021858087 + p * w; p = -0.00125372503 + p * w; p = -0.00417768164 + p * w; p = 0.246640727 + p * w;
        p = 1.50140941 + p * w;
    } else {
        w = std::sqrt(w) - 3.0;
        p = -0.000200214257; p = 0.000100950558 + p * w; p = 0.00134934322 + p * w; p = -0.00367342844 + p * w;
        p = 0.00573950773 + p * w; p = -0.0076224613 + p * w; p = 0.00943887047 + p * w; p = 1.00167406 + p * w;
        p = 2.83297682 + p * w;
    }
    double y = p * x;
    const double two_over_sqrtpi = 1.1283791670955126;
    for (int it = 0; it < 4; ++it) {
        double e;                                                // erf(y) - x, without cancellation in the tails
        if (std::fabs(x) < 0.5) e = std::erf(y) - x;
        else if (x > 0) e = (1.0 - x) - std::erfc(y);
        else e = std::erfc(-y) - (1.0 + x);
        double d = two_over_sqrtpi * std::exp(-y * y);
        y = y - e / (d + y * e);                                                       // Halley
    }
    return y;
}
static double random_normal1(Key k) {                                   // jax.random.normal(key, (1,)) float64
    uint64_t bits; random_bits64(k, 1, &bits);
    uint64_t fb = (bits >> 12) | 0x3FF0000000000000ull;
    double f; std::memcpy(&f, &fb, 8); f -= 1.0;
    const double lo = std::nextafter(-1.0, 0.0), hi = 1.0;
    double u = std::fmax(lo, f * (hi - lo) + lo);
    return std::sqrt(2.0) * erfinv_f64(u);
}

// =============================================================================================
// release model (main.py:209-280), templated so that its Jacobian (perturbative.py:281-296) is AD too
// =============================================================================================
template <class T> static inline T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> static inline void cross3(const T* a, const T* b, T* c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0]; }

template <class T>
static void release_T(const Program& P, double G, const T* x, const T* v, double Msat, double t,
                      const double* kv /*8*/, const double* nrm /*4 standard normals*/, T* out /*12: pos_lead,pos_trail,v_lead,v_trail*/) {
    // omega (main.py:88-96)
    T rad2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    T L[3]; cross3(x, v, L);
    T om[3] = {L[0] / rad2, L[1] / rad2, L[2] / rad2};
    T omega = sqrt(dot3(om, om));
    // d2phidr2 (main.py:76-85): r_hat is a closed-over constant inside the differentiated function
    T r = sqrt(rad2);
    T rhat[3] = {x[0] / r, x[1] / r, x[2] / r};
    T H[3][3]; hessian<T>(P, x, t, H);
    T d2 = T(0.0);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) d2 = d2 + rhat[i] * H[i][j] * rhat[j];
    // tidalr (main.py:104)
    T rt = powc((G * Msat) / (omega * omega - d2), 1.0 / 3.0);
    T relv = omega * rt;                 // main.py:238
    T vcirc = relv;                      // main.py:242
    T Lmag = sqrt(dot3(L, L));
    T zhat[3] = {L[0] / Lmag, L[1] / Lmag, L[2] / Lmag};
    T vr = dot3(v, rhat);
    T phiv[3] = {v[0] - vr * rhat[0], v[1] - vr * rhat[1], v[2] - vr * rhat[2]};
    T pn = sqrt(dot3(phiv, phiv));
    T phat[3] = {phiv[0] / pn, phiv[1] / pn, phiv[2] / pn};
    // main.py:263-266
    double kr = kv[0] + nrm[0] * kv[4];
    double kvphi = kr * (kv[1] + nrm[1] * kv[5]);
    double kz = kv[2] + nrm[2] * kv[6];
    double kvz = kv[3] + nrm[3] * kv[7];
    for (int c = 0; c < 3; ++c) {
        T pos_trail = x[c] + kr * rhat[c] * rt;                        // main.py:269
        pos_trail = pos_trail + zhat[c] * kz * (rt / 1.0);            // main.py:270
        T v_trail = v[c] + (0.0 + kvphi * vcirc * 1.0) * phat[c];     // main.py:271
        v_trail = v_trail + (kvz * vcirc * 1.0) * zhat[c];            // main.py:272
        T pos_lead = x[c] + kr * rhat[c] * (-rt);                      // main.py:275
        pos_lead = pos_lead + zhat[c] * kz * (-rt / 1.0);             // main.py:276
        T v_lead = v[c] + (0.0 + kvphi * vcirc * (-1.0)) * phat[c];   // main.py:277
        v_lead = v_lead + (kvz * vcirc * (-1.0)) * zhat[c];           // main.py:278
        out[c] = pos_lead; out[3 + c] = pos_trail; out[6 + c] = v_lead; out[9 + c] = v_trail;
    }
}

static void release_normals(int64_t seed, int64_t i, double* nrm4) {    // main.py:220-228, 263-266
    int64_t r[5]; randint5(prng_key(seed), 0, 1000, r);
    for (int q = 0; q < 4; ++q) nrm4[q] = random_normal1(prng_key(i * r[q]));
}

// =============================================================================================
// fields
// =============================================================================================
struct OrbitField {                    // Potential.velocity_acceleration (main.py:116-120)
    const Program* P;
    void operator()(double t, const double* y, double* dy) const {
        double g[3]; gradient<double>(*P, y, t, g);
        dy[0] = y[3]; dy[1] = y[4]; dy[2] = y[5]; dy[3] = -g[0]; dy[4] = -g[1]; dy[5] = -g[2];
    }
};

struct ResponseField {                 // MassRadiusPerturbation_OTF.term (fields.py:175-206)
    const Program* base; const SubhaloSet* S; SubhaloSet Sdr;   // Sdr: same set with dradius=1 (perturbative.py:681-690)
    int nsh;
    void operator()(double t, const double* y, double* dy) const {
        const double* x0 = y;
        double g[3]; gradient<double>(*base, x0, t, g);                         // fields.py:188
        double H[3][3]; hessian<double>(*base, x0, t, H);                       // fields.py:193 (d2H_dq2 = -H)
        dy[0] = y[3]; dy[1] = y[4]; dy[2] = y[5]; dy[3] = -g[0]; dy[4] = -g[1]; dy[5] = -g[2];
        typedef Dual<double, 3> D;
        D xd[3]; for (int i = 0; i < 3; ++i) xd[i] = D::var(x0[i], i);
        for (int j = 0; j < nsh; ++j) {
            const double* d = y + 6 + 12 * j; double* o = dy + 6 + 12 * j;
            D p1 = phi_subhalo<D>(*S, j, xd, t);                                // fields.py:191 (jacfwd of potential_per_SH)
            D p2 = phi_subhalo<D>(Sdr, j, xd, t);                               // fields.py:200
            for (int i = 0; i < 3; ++i) {
                double tx1 = 0, tdx = 0;
                for (int q = 0; q < 3; ++q) { tx1 += -H[i][q] * d[q]; tdx += -H[i][q] * d[6 + q]; }
                o[i] = d[3 + i];                                                // fields.py:195
                o[3 + i] = -p1.d[i] + tx1;                                      // fields.py:197
                o[6 + i] = d[9 + i];                                            // fields.py:201
                o[9 + i] = -p2.d[i] + tdx;                                      // fields.py:202
            }
        }
    }
};

struct ResponseField2 {                // MassRadiusPerturbation_OTF_SecondOrder.term (fields.py:289-320): y = [w(6), D(nsh,12), E(nsh,6)]
    const Program* base; const SubhaloSet* S; SubhaloSet Sdr; int nsh;
    void operator()(double t, const double* y, double* dy) const {
        const double* x0 = y;
        double g[3]; gradient<double>(*base, x0, t, g);
        typedef Dual<double, 3> D;
        D xd[3]; for (int i = 0; i < 3; ++i) xd[i] = D::var(x0[i], i);
        D Hd[3][3]; hessian<D>(*base, xd, t, Hd);                              // value = Hess, .d = third derivatives (fields.py:279-280)
        dy[0] = y[3]; dy[1] = y[4]; dy[2] = y[5]; dy[3] = -g[0]; dy[4] = -g[1]; dy[5] = -g[2];
        typedef Dual<D, 3> DD;
        const double* Dm = y + 6; const double* E = y + 6 + 12 * (size_t)nsh;
        double* dD = dy + 6; double* dE = dy + 6 + 12 * (size_t)nsh;
        for (int j = 0; j < nsh; ++j) {
            const double* d = Dm + 12 * j; const double* e = E + 6 * j;
            // per-subhalo gradient and Hessian of the perturbing potential (fields.py:282-283)
            DD xdd[3]; for (int i = 0; i < 3; ++i) xdd[i] = DD::var(xd[i], i);
            DD ph = phi_subhalo<DD>(*S, j, xdd, t);
            D p2 = phi_subhalo<D>(Sdr, j, xd, t);
            for (int i = 0; i < 3; ++i) {
                double tx1 = 0, tx2 = 0, tdx = 0, quad = 0, hp = 0;
                for (int q = 0; q < 3; ++q) {
                    tx1 += -Hd[i][q].v * d[q]; tx2 += -Hd[i][q].v * e[q]; tdx += -Hd[i][q].v * d[6 + q];
                    hp += -ph.d[i].d[q] * d[q];                                                       // dapert_dx . x1 (fields.py:312)
                    for (int r = 0; r < 3; ++r) quad += -Hd[i][q].d[r] * d[q] * d[r];                   // d2a_dx2[i,k,j] x1_k x1_j (fields.py:310-311)
                }
                dD[12 * j + i] = d[3 + i];
                dD[12 * j + 3 + i] = -ph.d[i].v + tx1;                                                // da (fields.py:307)
                dD[12 * j + 6 + i] = d[9 + i];
                dD[12 * j + 9 + i] = -p2.d[i] + tdx;                                                  // fields.py:314-315
                dE[6 * j + i] = e[3 + i];
                dE[6 * j + 3 + i] = tx2 + quad + hp;                                                  // d2a (fields.py:309-312)
            }
        }
    }
};

// =============================================================================================
// C ABI for the ctypes binding (oracle/__init__.py)
// =============================================================================================
extern "C" {

void* orc_program_new() { return new Program(); }
void orc_program_free(void* h) { delete (Program*)h; }
int orc_add_track(void* h, int kind, int n, const double* t, const double* y, int dim) {
    Program* P = (Program*)h; Track T; T.kind = kind; T.n = n; T.dim = dim;
    T.t.assign(t, t + n); T.y.assign(y, y + (size_t)n * dim); T.finalize();
    P->tracks.push_back(T); return (int)P->tracks.size() - 1;
}
// knot slopes of a cubic track supplied by the caller (any C1 piecewise cubic through the knots, e.g. a not-a-knot spline) instead of the
// interpax 'cubic' slopes Track::finalize computes
void orc_set_track_slopes(void* h, int track, const double* s) {
    Program* P = (Program*)h; Track& T = P->tracks[track];
    T.s.assign(s, s + (size_t)T.n * T.dim);
}
int orc_add_comp(void* h, int type, const double* p, int track) {
    Program* P = (Program*)h; Comp c; c.type = type; std::memcpy(c.p, p, sizeof(c.p)); c.track = track;
    P->comps.push_back(c); return (int)P->comps.size() - 1;
}
// GrowingPotential (potential.py:464-477): component `comp` is multiplied by the first column of track `track` evaluated at t
void orc_set_growth(void* h, int comp, int track) { Program* P = (Program*)h; P->comps[comp].growth = track; }
int orc_add_subhalos(void* h, int profile, int dradius, double G, int n, const double* m, const double* rs,
                     const double* x0, const double* v, const double* t0, const double* tw, int track) {
    Program* P = (Program*)h; SubhaloSet S; S.n = n; S.profile = profile; S.dradius = dradius; S.G = G;
    S.m.assign(m, m + n); S.rs.assign(rs, rs + n); S.x0.assign(x0, x0 + 3 * n); S.v.assign(v, v + 3 * n);
    S.t0.assign(t0, t0 + n); S.tw.assign(tw, tw + n);
    P->shs.push_back(S);
    Comp c; c.type = C_SUBHALOS; c.sh = (int)P->shs.size() - 1; c.track = track; P->comps.push_back(c);
    return c.sh;
}
int orc_num_threads() { int n = (int)std::thread::hardware_concurrency(); return n > 0 ? n : 1; }

void orc_potential(void* h, int n, const double* xyz, const double* t, double* out) {
    const Program& P = *(Program*)h;
    for (int i = 0; i < n; ++i) out[i] = phi_total<double>(P, xyz + 3 * i, t[i]);
}
void orc_gradient(void* h, int n, const double* xyz, const double* t, double* out) {
    const Program& P = *(Program*)h;
    for (int i = 0; i < n; ++i) gradient<double>(P, xyz + 3 * i, t[i], out + 3 * i);
}
void orc_hessian(void* h, int n, const double* xyz, const double* t, double* out) {
    const Program& P = *(Program*)h;
    for (int i = 0; i < n; ++i) { double H[3][3]; hessian<double>(P, xyz + 3 * i, t[i], H); std::memcpy(out + 9 * i, H, 72); }
}
void orc_third(void* h, int n, const double* xyz, const double* t, double* out /*[n,3,3,3]*/) {
    const Program& P = *(Program*)h;
    typedef Dual<double, 3> D;
    for (int i = 0; i < n; ++i) {
        D xd[3]; for (int c = 0; c < 3; ++c) xd[c] = D::var(xyz[3 * i + c], c);
        D H[3][3]; hessian<D>(P, xd, t[i], H);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int c = 0; c < 3; ++c) out[27 * i + 9 * a + 3 * b + c] = H[a][b].d[c];
    }
}
// per-subhalo potential values and x-gradients (potential_per_SH + jacfwd, perturbative.py:40-41, 695-696)
void orc_per_sh(void* h, int sh, const double* xyz, double t, double* phi /*[nsh]*/, double* grad /*[nsh,3]*/) {
    const Program& P = *(Program*)h; const SubhaloSet& S = P.shs[sh];
    typedef Dual<double, 3> D;
    D xd[3]; for (int c = 0; c < 3; ++c) xd[c] = D::var(xyz[c], c);
    for (int j = 0; j < S.n; ++j) { D p = phi_subhalo<D>(S, j, xd, t); phi[j] = p.v; for (int c = 0; c < 3; ++c) grad[3 * j + c] = p.d[c]; }
}
void orc_track_eval(void* h, int track, int n, const double* t, double* out, double* dout) {
    const Program& P = *(Program*)h; const Track& T = P.tracks[track];
    for (int i = 0; i < n; ++i) T.eval(t[i], out + (size_t)i * T.dim, dout ? dout + (size_t)i * T.dim : nullptr);
}

// batch of orbits: integrate_orbit / integrate_orbit_batch_* (main.py:125-202).  ts is [N,M] (per orbit).
void orc_integrate_orbits(void* h, int N, const double* w0, const double* t0, const double* t1, const double* ts, int M,
                          int solver, double rtol, double atol, double dtmin, double dtmax, int max_steps,
                          double* ys /*[N,M,6]*/, int* status /*[N]*/, int* nsteps /*[N,3] steps,acc,rej*/, int parallel /*threads*/) {
    const Program& P = *(Program*)h;
    Ctrl c; c.solver = solver; c.rtol = rtol; c.atol = atol; c.dtmin = dtmin; c.dtmax = dtmax; c.max_steps = max_steps;
    parallel_for(N, parallel, 16, [&](int i) {
        OrbitField f{&P};
        Stats s = solve(f, 6, t0[i], t1[i], w0 + 6 * (size_t)i, ts + (size_t)i * M, M, c, ys + (size_t)i * M * 6);
        status[i] = s.status; nsteps[3 * i] = s.n_steps; nsteps[3 * i + 1] = s.n_acc; nsteps[3 * i + 2] = s.n_rej;
    });
}

// one orbit with every accepted step recorded (for dense-output tests): returns number of accepted steps
int orc_orbit_steps(void* h, const double* w0, double t0, double t1, int solver, double rtol, double atol, double dtmin,
                    double dtmax, int max_steps, int cap, double* tgrid /*[cap+1]*/, double* ygrid /*[cap+1,6]*/) {
    const Program& P = *(Program*)h;
    Ctrl c; c.solver = solver; c.rtol = rtol; c.atol = atol; c.dtmin = dtmin; c.dtmax = dtmax; c.max_steps = max_steps;
    OrbitField f{&P};
    int cnt = 0;
    auto rec = [&](double ta, double tb, const double* y0, const double* y1, const double*) {
        if (cnt == 0) { tgrid[0] = ta; std::memcpy(ygrid, y0, 48); }
        if (cnt < cap) { tgrid[cnt + 1] = tb; std::memcpy(ygrid + 6 * (cnt + 1), y1, 48); }
        cnt++;
    };
    double dummy[6], tsd = t1;
    solve(f, 6, t0, t1, w0, &tsd, 1, c, dummy, rec);
    return cnt;
}

// one orbit, every step ATTEMPT traced: out[cap][4] = {tprev, dt, err, keep} in mirrored time; returns the number of attempts
int orc_orbit_trace(void* h, const double* w0, double t0, double t1, int solver, double rtol, double atol, double dtmin,
                    double dtmax, int max_steps, int cap, double* out, double* yfin /*[6]*/) {
    const Program& P = *(Program*)h;
    Ctrl c; c.solver = solver; c.rtol = rtol; c.atol = atol; c.dtmin = dtmin; c.dtmax = dtmax; c.max_steps = max_steps;
    std::vector<double> tr;
    c.trace = &tr;
    OrbitField f{&P};
    double tsd = t1;
    solve(f, 6, t0, t1, w0, &tsd, 1, c, yfin);
    const int n = (int)(tr.size() / 4);
    std::memcpy(out, tr.data(), sizeof(double) * 4 * (size_t)(n < cap ? n : cap));
    return n;
}

void orc_threefry(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t* out) { threefry2x32(k0, k1, c0, c1, out, out + 1); }
void orc_randint5(int64_t seed, int64_t lo, int64_t hi, int64_t* out) { randint5(prng_key(seed), lo, hi, out); }
double orc_normal1(int64_t seed) { return random_normal1(prng_key(seed)); }
double orc_erfinv(double x) { return erfinv_f64(x); }
void orc_release_normals(int64_t seed, int n, const int64_t* idx, double* out /*[n,4]*/) {
    for (int i = 0; i < n; ++i) release_normals(seed, idx[i], out + 4 * i);
}

// release_model over a batch (main.py:209-280 vmapped at main.py:293-303).  normals==NULL -> jax.random recipe.
void orc_release(void* h, double G, int n, const double* xv /*[n,6]*/, const double* Msat, const int64_t* idx, const double* t,
                 int64_t seed, const double* kvals /*8*/, const double* normals /*[n,4] or NULL*/, double* out /*[n,12]*/) {
    const Program& P = *(Program*)h;
    for (int i = 0; i < n; ++i) {
        double nr[4];
        if (normals) std::memcpy(nr, normals + 4 * i, 32); else release_normals(seed, idx[i], nr);
        release_T<double>(P, G, xv + 6 * i, xv + 6 * i + 3, Msat[i], t[i], kvals, nr, out + 12 * i);
    }
}
// release_model_Chen25 over a batch (streamhelpers.py:352-459); key words of the jax PRNG key; normals==NULL -> threefry recipe
void orc_release_chen25(void* h, double G, int n, const double* xv, const double* Msat, const double* t, uint32_t k0, uint32_t k1,
                        const double* mean, const double* factor, const double* normals /*[n,6] or NULL*/, double* out /*[n,12]*/) {
    const Program& P = *(Program*)h;
    for (int i = 0; i < n; ++i) {
        double z[6];
        if (normals) std::memcpy(z, normals + 6 * i, 48);
        else {
            uint32_t ki[2];
            for (int hh = 0; hh < 2; ++hh) {                   // jax.random.split(key, n)[i]
                const int f = 2 * i + hh, j = f < n ? f : f - n;
                uint32_t o0, o1; threefry2x32(k0, k1, (uint32_t)j, (uint32_t)(n + j), &o0, &o1);
                ki[hh] = f < n ? o0 : o1;
            }
            uint64_t bits[6]; random_bits64(Key{ki[0], ki[1]}, 6, bits);
            for (int q = 0; q < 6; ++q) {
                uint64_t fb = (bits[q] >> 12) | 0x3FF0000000000000ull; double f; std::memcpy(&f, &fb, 8); f -= 1.0;
                const double lo = std::nextafter(-1.0, 0.0);
                z[q] = std::sqrt(2.0) * erfinv_f64(std::fmax(lo, f * (1.0 - lo) + lo));
            }
        }
        double pv[6];
        for (int r = 0; r < 6; ++r) { pv[r] = mean[r]; for (int c = 0; c < 6; ++c) pv[r] += factor[6 * r + c] * z[c]; }
        const double* x = xv + 6 * i; const double* v = x + 3;
        double L[3]; cross3(x, v, L);
        const double rad2 = dot3(x, x), r = std::sqrt(rad2), Lm = std::sqrt(dot3(L, L)), omega = Lm / rad2;
        double xh[3] = {x[0] / r, x[1] / r, x[2] / r}, zh[3] = {L[0] / Lm, L[1] / Lm, L[2] / Lm};
        double H[3][3]; hessian<double>(P, x, t[i], H);
        double d2 = 0; for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) d2 += xh[a] * H[a][b] * xh[b];
        const double GM = G * Msat[i], rt = std::pow(GM / (omega * omega - d2), 1.0 / 3.0);          // main.py:104
        const double vr = dot3(v, xh);
        double ph[3] = {v[0] - vr * xh[0], v[1] - vr * xh[1], v[2] - vr * xh[2]};
        const double pn = std::sqrt(dot3(ph, ph));
        double yh[3] = {ph[0] / pn, ph[1] / pn, ph[2] / pn};
        const double Dr = pv[0] * rt, Dv = pv[3] * std::sqrt(2 * GM / Dr), d2r = 0.017453292519943295;           // streamhelpers.py:389-398
        const double phi = pv[1] * d2r, th = pv[2] * d2r, al = pv[4] * d2r, be = pv[5] * d2r;
        for (int k = 0; k < 3; ++k) {
            const double dx = (Dr * std::cos(th) * std::cos(phi)) * xh[k] + (Dr * std::cos(th) * std::sin(phi)) * yh[k];
            const double dvv = (Dv * std::cos(be) * std::cos(al)) * xh[k] + (Dv * std::cos(be) * std::sin(al)) * yh[k];
            out[12 * i + k] = x[k] - dx + (Dr * std::sin(th)) * zh[k];            // lead  (streamhelpers.py:419-430)
            out[12 * i + 3 + k] = x[k] + dx + (Dr * std::sin(th)) * zh[k];        // trail (streamhelpers.py:405-416)
            out[12 * i + 6 + k] = v[k] - dvv + (Dv * std::sin(be)) * zh[k];
            out[12 * i + 9 + k] = v[k] + dvv + (Dv * std::sin(be)) * zh[k];
        }
    }
}

// jacfwd(release_func) (perturbative.py:281-296): out[n,2,6,6], rows = (pos,vel) of lead / trail, cols = d/d(x,v)
void orc_release_jacobian(void* h, double G, int n, const double* xv, const double* Msat, const int64_t* idx, const double* t,
                          int64_t seed, const double* kvals, const double* normals, double* out) {
    const Program& P = *(Program*)h;
    typedef Dual<double, 6> D;
    for (int i = 0; i < n; ++i) {
        double nr[4];
        if (normals) std::memcpy(nr, normals + 4 * i, 32); else release_normals(seed, idx[i], nr);
        D x[3], v[3], o[12];
        for (int c = 0; c < 3; ++c) { x[c] = D::var(xv[6 * i + c], c); v[c] = D::var(xv[6 * i + 3 + c], 3 + c); }
        release_T<D>(P, G, x, v, Msat[i], t[i], kvals, nr, o);
        double* J = out + (size_t)i * 72;
        for (int c = 0; c < 3; ++c) for (int q = 0; q < 6; ++q) {
            J[0 * 36 + c * 6 + q] = o[c].d[q];            // lead pos
            J[0 * 36 + (3 + c) * 6 + q] = o[6 + c].d[q];  // lead vel
            J[1 * 36 + c * 6 + q] = o[3 + c].d[q];        // trail pos
            J[1 * 36 + (3 + c) * 6 + q] = o[9 + c].d[q];  // trail vel
        }
    }
}

// compute_perturbation_OTF (perturbative.py:101-135, 425-454, 726-755): per particle ONE coupled ODE with state
// [w(6), D(nsh,12)] sharing one controller; keep the final state.  sh = subhalo set index inside `hsh`.
void orc_linear_response(void* hbase, void* hsh, int sh, int N, const double* w0 /*[N,6]*/, const double* D0 /*[N,nsh,12] or NULL*/,
                         const double* t0 /*[N]*/, double t1, int solver, double rtol, double atol, double dtmin, double dtmax,
                         int max_steps, double* wout /*[N,6]*/, double* Dout /*[N,nsh,12]*/, int* status, int* nsteps, int parallel) {
    const Program& B = *(Program*)hbase; const SubhaloSet& S = ((Program*)hsh)->shs[sh];
    Ctrl c; c.solver = solver; c.rtol = rtol; c.atol = atol; c.dtmin = dtmin; c.dtmax = dtmax; c.max_steps = max_steps;
    const int nsh = S.n, n = 6 + 12 * nsh;
    parallel_for(N, parallel, 1, [&](int i) {
        ResponseField f; f.base = &B; f.S = &S; f.Sdr = S; f.Sdr.dradius = 1; f.nsh = nsh;
        std::vector<double> y0(n, 0.0), yf(n);
        std::memcpy(y0.data(), w0 + 6 * (size_t)i, 48);
        if (D0) std::memcpy(y0.data() + 6, D0 + (size_t)i * 12 * nsh, sizeof(double) * 12 * nsh);
        double tsv = t1;
        Stats s = solve(f, n, t0[i], t1, y0.data(), &tsv, 1, c, yf.data());
        std::memcpy(wout + 6 * (size_t)i, yf.data(), 48);
        std::memcpy(Dout + (size_t)i * 12 * nsh, yf.data() + 6, sizeof(double) * 12 * nsh);
        status[i] = s.status; nsteps[3 * i] = s.n_steps; nsteps[3 * i + 1] = s.n_acc; nsteps[3 * i + 2] = s.n_rej;
    });
}

// the same coupled field for ONE trajectory with SaveAt(ts): the backward progenitor response (perturbative.py:53-60)
void orc_linear_response_saveat(void* hbase, void* hsh, int sh, const double* w0, const double* D0, double t0, double t1, const double* ts, int M,
                                int solver, double rtol, double atol, double dtmin, double dtmax, int max_steps, double* ws /*[M,6]*/,
                                double* Ds /*[M,nsh,12]*/, int* status, int* nsteps) {
    const Program& B = *(Program*)hbase; const SubhaloSet& S = ((Program*)hsh)->shs[sh];
    Ctrl c; c.solver = solver; c.rtol = rtol; c.atol = atol; c.dtmin = dtmin; c.dtmax = dtmax; c.max_steps = max_steps;
    const int nsh = S.n, n = 6 + 12 * nsh;
    ResponseField f; f.base = &B; f.S = &S; f.Sdr = S; f.Sdr.dradius = 1; f.nsh = nsh;
    std::vector<double> y0(n, 0.0), ys((size_t)M * n);
    std::memcpy(y0.data(), w0, 48);
    if (D0) std::memcpy(y0.data() + 6, D0, sizeof(double) * 12 * nsh);
    Stats s = solve(f, n, t0, t1, y0.data(), ts, M, c, ys.data());
    for (int m = 0; m < M; ++m) {
        std::memcpy(ws + 6 * (size_t)m, ys.data() + (size_t)m * n, 48);
        std::memcpy(Ds + (size_t)m * 12 * nsh, ys.data() + (size_t)m * n + 6, sizeof(double) * 12 * nsh);
    }
    *status = s.status; nsteps[0] = s.n_steps; nsteps[1] = s.n_acc; nsteps[2] = s.n_rej;
}

// compute_perturbation_second_order_OTF (perturbative.py:757-772): per particle [w, D(nsh,12), E(nsh,6)] -> final state
void orc_second_order_response(void* hbase, void* hsh, int sh, int N, const double* w0, const double* D0, const double* E0, const double* t0,
                               double t1, int solver, double rtol, double atol, double dtmin, double dtmax, int max_steps, double* wout,
                               double* Dout, double* Eout, int* status, int* nsteps, int parallel) {
    const Program& B = *(Program*)hbase; const SubhaloSet& S = ((Program*)hsh)->shs[sh];
    Ctrl c; c.solver = solver; c.rtol = rtol; c.atol = atol; c.dtmin = dtmin; c.dtmax = dtmax; c.max_steps = max_steps;
    const int nsh = S.n, n = 6 + 18 * nsh;
    parallel_for(N, parallel, 1, [&](int i) {
        ResponseField2 f; f.base = &B; f.S = &S; f.Sdr = S; f.Sdr.dradius = 1; f.nsh = nsh;
        std::vector<double> y0(n, 0.0), yf(n);
        std::memcpy(y0.data(), w0 + 6 * (size_t)i, 48);
        if (D0) std::memcpy(y0.data() + 6, D0 + (size_t)i * 12 * nsh, sizeof(double) * 12 * nsh);
        if (E0) std::memcpy(y0.data() + 6 + 12 * nsh, E0 + (size_t)i * 6 * nsh, sizeof(double) * 6 * nsh);
        double tsv = t1;
        Stats s = solve(f, n, t0[i], t1, y0.data(), &tsv, 1, c, yf.data());
        std::memcpy(wout + 6 * (size_t)i, yf.data(), 48);
        std::memcpy(Dout + (size_t)i * 12 * nsh, yf.data() + 6, sizeof(double) * 12 * nsh);
        std::memcpy(Eout + (size_t)i * 6 * nsh, yf.data() + 6 + 12 * nsh, sizeof(double) * 6 * nsh);
        status[i] = s.status; nsteps[3 * i] = s.n_steps; nsteps[3 * i + 1] = s.n_acc; nsteps[3 * i + 2] = s.n_rej;
    });
}
void orc_second_order_term(void* hbase, void* hsh, int sh, double t, const double* y, double* dy) {
    const Program& B = *(Program*)hbase; const SubhaloSet& S = ((Program*)hsh)->shs[sh];
    ResponseField2 f; f.base = &B; f.S = &S; f.Sdr = S; f.Sdr.dradius = 1; f.nsh = S.n;
    f(t, y, dy);
}

// RHS of the response field at one state (fields.py:175-206), for unit tests
void orc_response_term(void* hbase, void* hsh, int sh, double t, const double* y, double* dy) {
    const Program& B = *(Program*)hbase; const SubhaloSet& S = ((Program*)hsh)->shs[sh];
    ResponseField f; f.base = &B; f.S = &S; f.Sdr = S; f.Sdr.dradius = 1; f.nsh = S.n;
    f(t, y, dy);
}

// RestrictedNbody_generator.term integrated by integrate_field (RestrictedNbody.py:93-106, 131): the WHOLE (N,6) tracer
// array is ONE diffrax state -> one step sequence, RMS error norm over all 6N components.  The field is the external
// potential + progenitor monopole on its track = one potential program with a translating component.
struct SharedOrbitField {
    const Program* P; int N;
    void operator()(double t, const double* y, double* dy) const {
        for (int i = 0; i < N; ++i) {
            const double* w = y + 6 * (size_t)i; double* d = dy + 6 * (size_t)i;
            double g[3]; gradient<double>(*P, w, t, g);
            d[0] = w[3]; d[1] = w[4]; d[2] = w[5]; d[3] = -g[0]; d[4] = -g[1]; d[5] = -g[2];
        }
    }
};
void orc_shared_step_orbits(void* h, int N, const double* w0, double t0, double t1, const double* ts, int M, int solver, double rtol, double atol,
                            double dtmin, double dtmax, int max_steps, double* ys /*[M,N,6]*/, int* status, int* nsteps) {
    const Program& P = *(Program*)h;
    Ctrl c; c.solver = solver; c.rtol = rtol; c.atol = atol; c.dtmin = dtmin; c.dtmax = dtmax; c.max_steps = max_steps;
    SharedOrbitField f{&P, N};
    Stats s = solve(f, 6 * N, t0, t1, w0, ts, M, c, ys);
    *status = s.status; nsteps[0] = s.n_steps; nsteps[1] = s.n_acc; nsteps[2] = s.n_rej;
}

// Nbody_field.term (fields.py:134-155): softened all-pairs self gravity + external potential, state (N,6) as ONE ODE.
// force_ij on i from j = G m_i m_j (x_i - x_j) / d^3 with d^2 = |x_i-x_j|^2 + eps^2; the reference sums forces[j, i] over j
// (axis 0) = sum_j G m_j m_i (x_j - x_i)/d^3, then divides by m_i.
struct NbodyField {
    const Program* P; int N; const double* m; double G, eps;
    void operator()(double t, const double* y, double* dy) const {
        for (int i = 0; i < N; ++i) {
            const double* w = y + 6 * (size_t)i; double* d = dy + 6 * (size_t)i;
            double acc[3] = {0, 0, 0};
            for (int j = 0; j < N; ++j) {
                if (j == i) continue;
                const double* wj = y + 6 * (size_t)j;
                double dx[3] = {wj[0] - w[0], wj[1] - w[1], wj[2] - w[2]};
                double d2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2] + eps * eps;
                double dd = std::sqrt(d2);
                double fmag = G * m[j] * m[i] / d2;
                for (int k = 0; k < 3; ++k) acc[k] += fmag * dx[k] / dd;
            }
            double g[3] = {0, 0, 0};
            if (P) gradient<double>(*P, w, t, g);
            d[0] = w[3]; d[1] = w[4]; d[2] = w[5];
            for (int k = 0; k < 3; ++k) d[3 + k] = acc[k] / m[i] - g[k];
        }
    }
};
void orc_nbody(void* h /*ext potential or NULL*/, int N, const double* masses, double G, double eps, const double* w0, double t0, double t1,
               const double* ts, int M, int solver, double rtol, double atol, double dtmin, double dtmax, int max_steps, double* ys /*[M,N,6]*/,
               int* status, int* nsteps) {
    Ctrl c; c.solver = solver; c.rtol = rtol; c.atol = atol; c.dtmin = dtmin; c.dtmax = dtmax; c.max_steps = max_steps;
    NbodyField f{(const Program*)h, N, masses, G, eps};
    Stats s = solve(f, 6 * N, t0, t1, w0, ts, M, c, ys);
    *status = s.status; nsteps[0] = s.n_steps; nsteps[1] = s.n_acc; nsteps[2] = s.n_rej;
}
void orc_nbody_term(void* h, int N, const double* masses, double G, double eps, double t, const double* y, double* dy) {
    NbodyField f{(const Program*)h, N, masses, G, eps};
    f(t, y, dy);
}

// Variational (tangent) equations along an unperturbed orbit: examples/higher_order_variationalEqn.ipynb cell 3
// (second_order_field.term wrapped in fields.CustomField, fields.py:362-377).  State [w(6), M(6,6) = dw/dw_init,
// M2(6,6,6) = d2w/dw_init^2] flattened row-major; da/dw = [-Hess Phi, 0], d2a/dw2 = -Phi_ijk on the position block.
struct VariationalFieldOrc {
    const Program* P; int order;
    int dim() const { return order == 2 ? 258 : 42; }
    void operator()(double t, const double* y, double* dy) const {
        typedef Dual<double, 3> D;
        D xd[3]; for (int c = 0; c < 3; ++c) xd[c] = D::var(y[c], c);
        D H[3][3]; hessian<D>(*P, xd, t, H);                       // H[a][b].v = Phi_ab, .d[c] = Phi_abc
        double g[3]; gradient<double>(*P, y, t, g);
        for (int a = 0; a < 3; ++a) { dy[a] = y[3 + a]; dy[3 + a] = -g[a]; }
        const double* M = y + 6; double* dM = dy + 6;
        for (int k = 0; k < 6; ++k)
            for (int a = 0; a < 3; ++a) {
                dM[6 * a + k] = M[6 * (a + 3) + k];
                double acc = 0; for (int j = 0; j < 3; ++j) acc += -H[a][j].v * M[6 * j + k];
                dM[6 * (a + 3) + k] = acc;
            }
        if (order < 2) return;
        const double* M2 = y + 42; double* dM2 = dy + 42;
        for (int k = 0; k < 6; ++k)
            for (int l = 0; l < 6; ++l)
                for (int a = 0; a < 3; ++a) {
                    dM2[36 * a + 6 * k + l] = M2[36 * (a + 3) + 6 * k + l];
                    double acc = 0;
                    for (int j = 0; j < 3; ++j) acc += -H[a][j].v * M2[36 * j + 6 * k + l];            // term1: da/dw . d2w
                    for (int j = 0; j < 3; ++j) for (int q = 0; q < 3; ++q) acc += -H[a][j].d[q] * M[6 * q + l] * M[6 * j + k];   // term2
                    dM2[36 * (a + 3) + 6 * k + l] = acc;
                }
    }
};
void orc_variational(void* h, int order, int N, const double* w0, const double* M0 /*[N,6,6] or NULL = I*/, const double* M20 /*or NULL = 0*/,
                     const double* t0, double t1, int solver, double rtol, double atol, double dtmin, double dtmax, int max_steps, double* wout,
                     double* Mout, double* M2out, int* status, int* nsteps, int parallel) {
    const Program& P = *(Program*)h;
    Ctrl c; c.solver = solver; c.rtol = rtol; c.atol = atol; c.dtmin = dtmin; c.dtmax = dtmax; c.max_steps = max_steps;
    VariationalFieldOrc f{&P, order};
    const int n = f.dim();
    parallel_for(N, parallel, 1, [&](int i) {
        std::vector<double> y0(n, 0.0), yf(n);
        std::memcpy(y0.data(), w0 + 6 * (size_t)i, 48);
        if (M0) std::memcpy(y0.data() + 6, M0 + 36 * (size_t)i, 36 * 8); else for (int a = 0; a < 6; ++a) y0[6 + 7 * a] = 1.0;
        if (order == 2 && M20) std::memcpy(y0.data() + 42, M20 + 216 * (size_t)i, 216 * 8);
        double tsv = t1;
        Stats s = solve(f, n, t0[i], t1, y0.data(), &tsv, 1, c, yf.data());
        std::memcpy(wout + 6 * (size_t)i, yf.data(), 48);
        std::memcpy(Mout + 36 * (size_t)i, yf.data() + 6, 36 * 8);
        if (order == 2) std::memcpy(M2out + 216 * (size_t)i, yf.data() + 42, 216 * 8);
        status[i] = s.status; nsteps[3 * i] = s.n_steps; nsteps[3 * i + 1] = s.n_acc; nsteps[3 * i + 2] = s.n_rej;
    });
}
void orc_variational_term(void* h, int order, double t, const double* y, double* dy) {
    VariationalFieldOrc f{(const Program*)h, order};
    f(t, y, dy);
}

}  // extern "C"
