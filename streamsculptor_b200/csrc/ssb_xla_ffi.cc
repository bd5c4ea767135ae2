// XLA FFI shim for libssb200: the handlers `jax.ffi.ffi_call` binds (streamsculptor_b200/jax_plugin.py registers them), so that the
// reference's hot-path calls stay inside `jax.jit` (main.py:139-162, 343-368; perturbative.py:726-755).
//
// Built only where the XLA FFI headers exist (jaxlib ships them under jaxlib/include):
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jaxlib,os;print(os.path.join(os.path.dirname(jaxlib.__file__),'include'))") \
//       -I include streamsculptor_b200/csrc/ssb_xla_ffi.cc -L streamsculptor_b200/_lib -lssb200 -o libssb200_ffi.so
// This repository's image has neither jax nor those headers, so the file is compile-CHECKED against a stub of the public API
// (tools/xla_ffi_stub, `__graft_entry__.build()` and tests/test_host_cpu.py run g++ -fsyntax-only); it has never run under XLA.
//
// How a potential travels.  Attributes are compile-time constants under jax.jit, device pointers are not, so the potential program is
// split: the pointer-free part (components, parameters, table kinds / lengths, subhalo-set sizes) is the byte attribute `program`
// (struct ssb_ffi_program below, built by jax_plugin.flatten_program), and every table is an ordinary buffer operand that follows the
// handler's fixed operands - per track: t[n], y[n,3], s[n,3] (knot slopes; ignored for linear tracks); per subhalo set: m[n], r_s[n],
// x0[n,3], v[n,3], t0[n], t_window[n].  The handler assembles the ssb_potential that points at them.  Time-dependent potentials
// (moving perturbers, subhalo ensembles) therefore work with traced arrays, which the previous attribute-only design could not carry.
// Packed perturber sets (ssb_perturbers) are not carried yet: express them as translating components (<= SSB_MAX_TRACK) under jax.
#if __has_include("xla/ffi/api/ffi.h")
#include <cstring>

#include "../../include/ssb200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

struct ssb_ffi_program {
    int32_t n_comp, n_track, n_sh, _pad;
    ssb_component comp[SSB_MAX_COMP];
    int32_t track_kind[SSB_MAX_TRACK], track_n[SSB_MAX_TRACK];
    int32_t sh_n[SSB_MAX_SUBHALO_SETS], sh_profile[SSB_MAX_SUBHALO_SETS];
    double sh_G[SSB_MAX_SUBHALO_SETS];
};

static_assert(sizeof(ssb_ffi_program) == 16 + sizeof(ssb_component) * SSB_MAX_COMP + 8 * SSB_MAX_TRACK + 16 * SSB_MAX_SUBHALO_SETS,
              "ssb_ffi_program must stay padding-free: jax_plugin.FfiProgram mirrors it field by field");

// number of trailing operands a program needs
static size_t program_operands(const ssb_ffi_program& p) { return 3 * (size_t)p.n_track + 6 * (size_t)p.n_sh; }

// program attribute + trailing operands [first, first + program_operands) -> ssb_potential
static ffi::Error assemble(ffi::Span<const uint8_t> attr, ffi::RemainingArgs rest, size_t first, ssb_potential* pot, size_t* used) {
    if (attr.size() != sizeof(ssb_ffi_program)) return ffi::Error::InvalidArgument("potential program attribute has the wrong size");
    ssb_ffi_program p;
    std::memcpy(&p, attr.data(), sizeof(p));
    if (p.n_comp < 0 || p.n_comp > SSB_MAX_COMP || p.n_track < 0 || p.n_track > SSB_MAX_TRACK || p.n_sh < 0 || p.n_sh > SSB_MAX_SUBHALO_SETS)
        return ffi::Error::InvalidArgument("potential program: component / track / subhalo-set count out of range");
    if (first + program_operands(p) > rest.size()) return ffi::Error::InvalidArgument("potential program: table operands missing");
    std::memset(pot, 0, sizeof(*pot));
    pot->n_comp = p.n_comp; pot->n_track = p.n_track; pot->n_sh = p.n_sh;
    std::memcpy(pot->comp, p.comp, sizeof(p.comp));
    size_t k = first;
    auto f64 = [&](size_t i, const double** out, size_t want) -> bool {
        auto b = rest.get<ffi::Buffer<ffi::F64>>(i);
        if (!b.has_value() || b->element_count() < want) return false;
        *out = b->typed_data();
        return true;
    };
    for (int i = 0; i < p.n_track; ++i) {
        ssb_track& t = pot->track[i];
        t.kind = p.track_kind[i]; t.n = p.track_n[i];
        const size_t n = (size_t)t.n;
        if (!f64(k, &t.t, n) || !f64(k + 1, &t.y, 3 * n) || !f64(k + 2, &t.s, t.kind == SSB_TRACK_CUBIC ? 3 * n : 0))
            return ffi::Error::InvalidArgument("potential program: track operand has the wrong dtype / size");
        if (t.kind != SSB_TRACK_CUBIC) t.s = nullptr;
        k += 3;
    }
    for (int i = 0; i < p.n_sh; ++i) {
        ssb_subhalos& s = pot->sh[i];
        s.n = p.sh_n[i]; s.profile = p.sh_profile[i]; s.G = p.sh_G[i];
        const size_t n = (size_t)s.n;
        if (!f64(k, &s.m, n) || !f64(k + 1, &s.rs, n) || !f64(k + 2, &s.x0, 3 * n) || !f64(k + 3, &s.v, 3 * n) || !f64(k + 4, &s.t0, n) || !f64(k + 5, &s.tw, n))
            return ffi::Error::InvalidArgument("potential program: subhalo operand has the wrong dtype / size");
        k += 6;
    }
    *used = k - first;
    return ffi::Error::Success();
}

static ffi::Error status_of(int rc) { return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(ssb_last_error()); }

// ---- Potential.integrate_orbit / integrate_orbit_batch_vmapped (main.py:125-202) ----
static ffi::Error OrbitIntegrateImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> w0, ffi::Buffer<ffi::F64> t0, ffi::Buffer<ffi::F64> t1,
                                     ffi::Buffer<ffi::F64> ts, ffi::RemainingArgs tables, ffi::Span<const uint8_t> program, int32_t solver,
                                     int32_t max_steps, double rtol, double atol, double dtmin, double dtmax, ffi::ResultBuffer<ffi::F64> ys,
                                     ffi::ResultBuffer<ffi::S32> status, ffi::ResultBuffer<ffi::S32> nsteps) {
    ssb_potential pot;
    size_t used = 0;
    ffi::Error e = assemble(program, tables, 0, &pot, &used);
    if (!e.success()) return e;
    const int64_t N = w0.dimensions()[0];
    const bool per_orbit = ts.dimensions().size() == 2;
    const int32_t M = (int32_t)ts.dimensions()[ts.dimensions().size() - 1];
    ssb_ctrl c{solver, max_steps, rtol, atol, dtmin, dtmax};
    return status_of(ssb_orbit_integrate_f64(&pot, N, w0.typed_data(), t0.typed_data(), t1.typed_data(), ts.typed_data(), M, per_orbit, c,
                                             ys->typed_data(), status->typed_data(), nsteps->typed_data(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(ssb_orbit_integrate_ffi, OrbitIntegrateImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // w0 [N,6]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // t0 [N]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // t1 [N]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // ts [M] or [N,M]
                                  .RemainingArgs()                // tables of the potential program
                                  .Attr<ffi::Span<const uint8_t>>("program")
                                  .Attr<int32_t>("solver")
                                  .Attr<int32_t>("max_steps")
                                  .Attr<double>("rtol")
                                  .Attr<double>("atol")
                                  .Attr<double>("dtmin")
                                  .Attr<double>("dtmax")
                                  .Ret<ffi::Buffer<ffi::F64>>()   // ys [N,M,6]
                                  .Ret<ffi::Buffer<ffi::S32>>()   // status [N]
                                  .Ret<ffi::Buffer<ffi::S32>>()); // nsteps [N,3]

// ---- forward-mode derivative of integrate_orbit with respect to w0: the state-transition matrix (custom_jvp rule; main.py:160) ----
static ffi::Error VariationalImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> w0, ffi::Buffer<ffi::F64> t0, ffi::RemainingArgs tables,
                                  ffi::Span<const uint8_t> program, double t1, int32_t solver, int32_t max_steps, double rtol, double atol, double dtmin,
                                  double dtmax, ffi::ResultBuffer<ffi::F64> wout, ffi::ResultBuffer<ffi::F64> Mout, ffi::ResultBuffer<ffi::S32> status,
                                  ffi::ResultBuffer<ffi::S32> nsteps) {
    ssb_potential pot;
    size_t used = 0;
    ffi::Error e = assemble(program, tables, 0, &pot, &used);
    if (!e.success()) return e;
    ssb_ctrl c{solver, max_steps, rtol, atol, dtmin, dtmax};
    return status_of(ssb_variational_f64(&pot, 1, w0.dimensions()[0], w0.typed_data(), nullptr, nullptr, t0.typed_data(), t1, c, wout->typed_data(),
                                         Mout->typed_data(), nullptr, status->typed_data(), nsteps->typed_data(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(ssb_variational_ffi, VariationalImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // w0 [N,6]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // t0 [N]
                                  .RemainingArgs()
                                  .Attr<ffi::Span<const uint8_t>>("program")
                                  .Attr<double>("t1")
                                  .Attr<int32_t>("solver")
                                  .Attr<int32_t>("max_steps")
                                  .Attr<double>("rtol")
                                  .Attr<double>("atol")
                                  .Attr<double>("dtmin")
                                  .Attr<double>("dtmax")
                                  .Ret<ffi::Buffer<ffi::F64>>()   // w [N,6]
                                  .Ret<ffi::Buffer<ffi::F64>>()   // M [N,6,6] = d w / d w0
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

// ---- gen_stream_vmapped (main.py:343-368) as ONE custom call: progenitor orbit, release, 2(Nts-1) orbit solves.  `scratch` is an extra
// result buffer of ssb_stream_scratch_bytes(Nts, max_steps) bytes that XLA allocates (the library never allocates device memory).
// Trailing operands: the tables of `program`, then those of `program_release`. ----
static ffi::Error GenStreamImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> ts, ffi::Buffer<ffi::F64> prog_w0, ffi::Buffer<ffi::F64> msat,
                                ffi::RemainingArgs tables, ffi::Span<const uint8_t> program, ffi::Span<const uint8_t> program_release,
                                ffi::Span<const double> kvals, double G, int64_t seed, int32_t solver, int32_t max_steps, double rtol, double atol,
                                double dtmin, double dtmax, ffi::ResultBuffer<ffi::F64> lead_trail /*[2,Nts-1,6]*/,
                                ffi::ResultBuffer<ffi::S32> status /*[2,Nts-1]*/, ffi::ResultBuffer<ffi::S32> nsteps /*[2,Nts-1,3]*/,
                                ffi::ResultBuffer<ffi::U8> scratch) {
    if (kvals.size() != 8) return ffi::Error::InvalidArgument("gen_stream: kvals needs 8 values");
    ssb_potential pot, rel;
    size_t used = 0, used2 = 0;
    ffi::Error e = assemble(program, tables, 0, &pot, &used);
    if (!e.success()) return e;
    e = assemble(program_release, tables, used, &rel, &used2);
    if (!e.success()) return e;
    const int64_t Nts = ts.dimensions()[0], n = Nts - 1;
    ssb_ctrl c{solver, max_steps, rtol, atol, dtmin, dtmax};
    double* lt = lead_trail->typed_data();
    return status_of(ssb_gen_stream_f64(&pot, &rel, G, Nts, ts.typed_data(), prog_w0.typed_data(), msat.typed_data(), seed, kvals.data(), nullptr, c, 0, 1, n,
                                        lt, lt + 6 * n, status->typed_data(), nsteps->typed_data(), scratch->typed_data(), scratch->size_bytes(), stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(ssb_gen_stream_ffi, GenStreamImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // ts [Nts]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // prog_w0 [6]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // Msat [Nts]
                                  .RemainingArgs()
                                  .Attr<ffi::Span<const uint8_t>>("program")
                                  .Attr<ffi::Span<const uint8_t>>("program_release")
                                  .Attr<ffi::Span<const double>>("kvals")
                                  .Attr<double>("G")
                                  .Attr<int64_t>("seed")
                                  .Attr<int32_t>("solver")
                                  .Attr<int32_t>("max_steps")
                                  .Attr<double>("rtol")
                                  .Attr<double>("atol")
                                  .Attr<double>("dtmin")
                                  .Attr<double>("dtmax")
                                  .Ret<ffi::Buffer<ffi::F64>>()   // [2, Nts-1, 6] lead, trail
                                  .Ret<ffi::Buffer<ffi::S32>>()   // status
                                  .Ret<ffi::Buffer<ffi::S32>>()   // nsteps
                                  .Ret<ffi::Buffer<ffi::U8>>());  // scratch

// ---- compute_perturbation_OTF (perturbative.py:726-755) as one custom call.  The perturbing subhalo set is the LAST set of operands
// (six arrays after the base program's tables); scratch: ssb_response_scratch_bytes(n_sh) bytes. ----
static ffi::Error LinearResponseImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> w0, ffi::Buffer<ffi::F64> t0, ffi::RemainingArgs tables,
                                     ffi::Span<const uint8_t> program, int32_t n_sh, int32_t profile, double G, double t1, int32_t solver,
                                     int32_t max_steps, double rtol, double atol, double dtmin, double dtmax, ffi::ResultBuffer<ffi::F64> wout,
                                     ffi::ResultBuffer<ffi::F64> Dout, ffi::ResultBuffer<ffi::S32> status, ffi::ResultBuffer<ffi::S32> nsteps,
                                     ffi::ResultBuffer<ffi::U8> scratch) {
    ssb_potential pot;
    size_t used = 0;
    ffi::Error e = assemble(program, tables, 0, &pot, &used);
    if (!e.success()) return e;
    if (used + 6 > tables.size()) return ffi::Error::InvalidArgument("linear_response: subhalo operands missing");
    ssb_subhalos sh;
    std::memset(&sh, 0, sizeof(sh));
    sh.n = n_sh; sh.profile = profile; sh.G = G;
    const double** dst[6] = {&sh.m, &sh.rs, &sh.x0, &sh.v, &sh.t0, &sh.tw};
    const size_t want[6] = {1, 1, 3, 3, 1, 1};
    for (int i = 0; i < 6; ++i) {
        auto b = tables.get<ffi::Buffer<ffi::F64>>(used + i);
        if (!b.has_value() || b->element_count() < want[i] * (size_t)n_sh) return ffi::Error::InvalidArgument("linear_response: bad subhalo operand");
        *dst[i] = b->typed_data();
    }
    ssb_ctrl c{solver, max_steps, rtol, atol, dtmin, dtmax};
    return status_of(ssb_linear_response_f64(&pot, &sh, w0.dimensions()[0], w0.typed_data(), nullptr, t0.typed_data(), t1, c, wout->typed_data(),
                                             Dout->typed_data(), status->typed_data(), nsteps->typed_data(), scratch->typed_data(), scratch->size_bytes(),
                                             stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(ssb_linear_response_ffi, LinearResponseImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // w0 [N,6]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // t0 [N]
                                  .RemainingArgs()                // tables of the base program, then m, r_s, x0, v, t0, t_window of the subhalos
                                  .Attr<ffi::Span<const uint8_t>>("program")
                                  .Attr<int32_t>("n_sh")
                                  .Attr<int32_t>("profile")
                                  .Attr<double>("G")
                                  .Attr<double>("t1")
                                  .Attr<int32_t>("solver")
                                  .Attr<int32_t>("max_steps")
                                  .Attr<double>("rtol")
                                  .Attr<double>("atol")
                                  .Attr<double>("dtmin")
                                  .Attr<double>("dtmax")
                                  .Ret<ffi::Buffer<ffi::F64>>()   // w [N,6]
                                  .Ret<ffi::Buffer<ffi::F64>>()   // D [N,n_sh,12]
                                  .Ret<ffi::Buffer<ffi::S32>>()   // status [N]
                                  .Ret<ffi::Buffer<ffi::S32>>()   // nsteps [N,3]
                                  .Ret<ffi::Buffer<ffi::U8>>());  // scratch
#endif
