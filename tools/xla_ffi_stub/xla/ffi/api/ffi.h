// COMPILE-CHECK STUB of the public XLA FFI C++ API (xla/ffi/api/ffi.h as shipped in jaxlib/include), written from its documented
// surface so that streamsculptor_b200/csrc/ssb_xla_ffi.cc can be parsed and type-checked in an image without jax / jaxlib:
//   g++ -std=c++17 -fsyntax-only -I tools/xla_ffi_stub -I include streamsculptor_b200/csrc/ssb_xla_ffi.cc
// It declares only what the shim uses (Buffer / ResultBuffer / AnyBuffer / RemainingArgs / Span / Error / Ffi::Bind() builder /
// PlatformStream / XLA_FFI_DEFINE_HANDLER_SYMBOL) and does nothing at run time.  A real build uses jaxlib's header instead.
#ifndef SSB_XLA_FFI_STUB_H
#define SSB_XLA_FFI_STUB_H
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#define SSB_XLA_FFI_STUB 1
typedef struct CUstream_st* cudaStream_t;

namespace xla {
namespace ffi {

enum DataType { U8, S32, S64, F64 };
template <DataType> struct NativeType;
template <> struct NativeType<U8> { typedef uint8_t type; };
template <> struct NativeType<S32> { typedef int32_t type; };
template <> struct NativeType<S64> { typedef int64_t type; };
template <> struct NativeType<F64> { typedef double type; };

template <typename T>
class Span {
 public:
    Span() : d_(nullptr), n_(0) {}
    Span(const T* d, size_t n) : d_(d), n_(n) {}
    const T* data() const { return d_; }
    size_t size() const { return n_; }
    const T& operator[](size_t i) const { return d_[i]; }
    const T* begin() const { return d_; }
    const T* end() const { return d_ + n_; }
 private:
    const T* d_; size_t n_;
};

class Error {
 public:
    static Error Success() { return Error(); }
    static Error InvalidArgument(std::string m) { return Error(std::move(m)); }
    static Error Internal(std::string m) { return Error(std::move(m)); }
    bool success() const { return ok_; }
 private:
    Error() : ok_(true) {}
    explicit Error(std::string m) : ok_(false), msg_(std::move(m)) {}
    bool ok_; std::string msg_;
};

template <DataType dtype>
class Buffer {
 public:
    typedef typename NativeType<dtype>::type T;
    T* typed_data() const { return nullptr; }
    Span<int64_t> dimensions() const { return Span<int64_t>(); }
    size_t element_count() const { return 0; }
    size_t size_bytes() const { return 0; }
};
class AnyBuffer {
 public:
    void* untyped_data() const { return nullptr; }
    Span<int64_t> dimensions() const { return Span<int64_t>(); }
    size_t element_count() const { return 0; }
    size_t size_bytes() const { return 0; }
    DataType element_type() const { return F64; }
};
template <typename B> class Result { public: B* operator->() { return &b_; } B& operator*() { return b_; } private: B b_; };
template <DataType dtype> using ResultBuffer = Result<Buffer<dtype>>;

template <typename T> class ErrorOr { public: bool has_value() const { return true; } T& value() { return v_; } T* operator->() { return &v_; } private: T v_; };
class RemainingArgs {
 public:
    size_t size() const { return 0; }
    template <typename T> ErrorOr<T> get(size_t) const { return ErrorOr<T>(); }
};

template <typename T> struct PlatformStream {};

struct Binding {
    template <typename T> Binding& Ctx() { return *this; }
    template <typename T> Binding& Arg() { return *this; }
    template <typename T> Binding& Ret() { return *this; }
    template <typename T> Binding& Attr(const char*) { return *this; }
    Binding& RemainingArgs() { return *this; }
};
struct Ffi { static Binding Bind() { return Binding(); } };

}  // namespace ffi
}  // namespace xla

// the real macro instantiates a handler that decodes the call frame according to `binding` and calls `fn`; the stub only makes sure
// that `fn` and the binding expression compile
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, fn, binding)                 \
    extern "C" void* name(void* call_frame) {                            \
        (void)(binding); (void)&fn; (void)call_frame;                    \
        return nullptr;                                                  \
    }
#endif
