// XLA FFI shim for libssb200 (compiled only where the XLA FFI headers exist: `pip show jaxlib` ships them under
// jaxlib/include).  NOT built in this repository's image (no jax, no xla/ffi headers on disk) - see INTEGRATION.md.
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jaxlib,os;print(os.path.join(os.path.dirname(jaxlib.__file__),'include'))") \
//       -I include ssb_xla_ffi.cc -L streamsculptor_b200/_lib -lssb200 -o libssb200_ffi.so
#if __has_include("xla/ffi/api/ffi.h")
#include <cstring>

#include "../../include/ssb200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

// The potential program travels as an opaque byte attribute (sizeof(ssb_potential) bytes built on the Python side from the
// Potential object tree); its table / subhalo pointers are patched from the trailing buffer operands.
static ffi::Error OrbitIntegrateImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> w0, ffi::Buffer<ffi::F64> t0, ffi::Buffer<ffi::F64> t1,
                                     ffi::Buffer<ffi::F64> ts, ffi::Span<const uint8_t> program, int32_t solver, int32_t max_steps,
                                     double rtol, double atol, double dtmin, double dtmax, ffi::ResultBuffer<ffi::F64> ys,
                                     ffi::ResultBuffer<ffi::S32> status, ffi::ResultBuffer<ffi::S32> nsteps) {
    if (program.size() != sizeof(ssb_potential)) return ffi::Error::InvalidArgument("potential program has the wrong size");
    ssb_potential pot;
    std::memcpy(&pot, program.data(), sizeof(pot));
    const int64_t N = w0.dimensions()[0];
    const bool per_orbit = ts.dimensions().size() == 2;
    const int32_t M = (int32_t)ts.dimensions().back();
    ssb_ctrl c{solver, max_steps, rtol, atol, dtmin, dtmax};
    const int rc = ssb_orbit_integrate_f64(&pot, N, w0.typed_data(), t0.typed_data(), t1.typed_data(), ts.typed_data(), M, per_orbit, c,
                                           ys->typed_data(), status->typed_data(), nsteps->typed_data(), stream);
    return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(ssb_last_error());
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(ssb_orbit_integrate_ffi, OrbitIntegrateImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // w0 [N,6]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // t0 [N]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // t1 [N]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // ts [M] or [N,M]
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

// gen_stream_vmapped (main.py:343-368) as ONE custom call: progenitor orbit, release, 2(Nts-1) orbit solves.  `scratch` is an extra
// result buffer of ssb_stream_scratch_bytes(Nts, max_steps) bytes that XLA allocates (the library never allocates device memory).
static ffi::Error GenStreamImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> ts, ffi::Buffer<ffi::F64> prog_w0, ffi::Buffer<ffi::F64> msat,
                                ffi::Span<const uint8_t> program, ffi::Span<const uint8_t> program_release, ffi::Span<const double> kvals, double G,
                                int64_t seed, int32_t solver, int32_t max_steps, double rtol, double atol, double dtmin, double dtmax,
                                ffi::ResultBuffer<ffi::F64> lead_trail /*[2,Nts-1,6]*/, ffi::ResultBuffer<ffi::S32> status /*[2,Nts-1]*/,
                                ffi::ResultBuffer<ffi::S32> nsteps /*[2,Nts-1,3]*/, ffi::ResultBuffer<ffi::U8> scratch) {
    if (program.size() != sizeof(ssb_potential) || program_release.size() != sizeof(ssb_potential) || kvals.size() != 8)
        return ffi::Error::InvalidArgument("gen_stream: bad program / kvals attribute");
    ssb_potential pot, rel;
    std::memcpy(&pot, program.data(), sizeof(pot));
    std::memcpy(&rel, program_release.data(), sizeof(rel));
    const int64_t Nts = ts.dimensions()[0], n = Nts - 1;
    ssb_ctrl c{solver, max_steps, rtol, atol, dtmin, dtmax};
    double* lt = lead_trail->typed_data();
    const int rc = ssb_gen_stream_f64(&pot, &rel, G, Nts, ts.typed_data(), prog_w0.typed_data(), msat.typed_data(), seed, kvals.data(), nullptr, c, 0, 1, n,
                                      lt, lt + 6 * n, status->typed_data(), nsteps->typed_data(), scratch->typed_data(), scratch->size_bytes(), stream);
    return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(ssb_last_error());
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(ssb_gen_stream_ffi, GenStreamImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // ts [Nts]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // prog_w0 [6]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // Msat [Nts]
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

// compute_perturbation_OTF (perturbative.py:726-755) as one custom call.  The subhalo arrays are ordinary operands; the struct that
// points at them is assembled here.  scratch: ssb_response_scratch_bytes(n_sh) bytes.
static ffi::Error LinearResponseImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> w0, ffi::Buffer<ffi::F64> t0, ffi::Buffer<ffi::F64> m,
                                     ffi::Buffer<ffi::F64> rs, ffi::Buffer<ffi::F64> x0, ffi::Buffer<ffi::F64> v, ffi::Buffer<ffi::F64> sh_t0,
                                     ffi::Buffer<ffi::F64> tw, ffi::Span<const uint8_t> program, int32_t profile, double G, double t1, int32_t solver,
                                     int32_t max_steps, double rtol, double atol, double dtmin, double dtmax, ffi::ResultBuffer<ffi::F64> wout,
                                     ffi::ResultBuffer<ffi::F64> Dout, ffi::ResultBuffer<ffi::S32> status, ffi::ResultBuffer<ffi::S32> nsteps,
                                     ffi::ResultBuffer<ffi::U8> scratch) {
    if (program.size() != sizeof(ssb_potential)) return ffi::Error::InvalidArgument("linear_response: potential program has the wrong size");
    ssb_potential pot;
    std::memcpy(&pot, program.data(), sizeof(pot));
    ssb_subhalos sh;
    sh.n = (int32_t)m.dimensions()[0]; sh.profile = profile; sh.G = G;
    sh.m = m.typed_data(); sh.rs = rs.typed_data(); sh.x0 = x0.typed_data(); sh.v = v.typed_data(); sh.t0 = sh_t0.typed_data(); sh.tw = tw.typed_data();
    ssb_ctrl c{solver, max_steps, rtol, atol, dtmin, dtmax};
    const int rc = ssb_linear_response_f64(&pot, &sh, w0.dimensions()[0], w0.typed_data(), nullptr, t0.typed_data(), t1, c, wout->typed_data(),
                                           Dout->typed_data(), status->typed_data(), nsteps->typed_data(), scratch->typed_data(), scratch->size_bytes(),
                                           stream);
    return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(ssb_last_error());
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(ssb_linear_response_ffi, LinearResponseImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // w0 [N,6]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // t0 [N]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // m [n_sh]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // r_s [n_sh]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // x0 [n_sh,3]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // v [n_sh,3]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // t0 [n_sh]
                                  .Arg<ffi::Buffer<ffi::F64>>()   // t_window [n_sh]
                                  .Attr<ffi::Span<const uint8_t>>("program")
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
