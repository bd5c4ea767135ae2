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
#endif
