// TEST DOUBLE of the XLA FFI C++ binding header (xla/ffi/api/ffi.h, shipped by jaxlib under jax.ffi.include_dir()).
//
// jaxlib is not installable in the build image (no network), so waveforminversionust_b200/csrc/xla_ffi_shim.cc could not
// meet a compiler.  This header re-declares, with the same names and call shapes, the small part of that API the shim
// uses -- Buffer<dtype> / ResultBuffer<dtype>, Error, PlatformStream, Ffi::Bind().Ctx().Arg().Attr().Ret() and
// XLA_FFI_DEFINE_HANDLER_SYMBOL -- over a plain call frame that a test can fill from Python (stub_frame.cc), so that the
// shim's own logic (plan cache, host copies of the small operands, forwarding of XLA's stream and device pointers to the
// C ABI, error conversion) is compiled and exercised on the GPU.  It is NOT the XLA ABI: a library built against this
// header cannot be registered with jax; build against the real header for that (jax_frontend.build_ffi()).
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace xla {
namespace ffi {

enum class DataType : int { F32 = 11, F64 = 12, C64 = 15, S32 = 4 };
inline constexpr DataType F32 = DataType::F32;
inline constexpr DataType F64 = DataType::F64;
inline constexpr DataType C64 = DataType::C64;
inline constexpr DataType S32 = DataType::S32;

namespace stub {
struct RawBuffer { void* data; int dtype; std::vector<int64_t> dims; };
struct CallFrame {
    void* stream = nullptr;
    std::vector<RawBuffer> args, rets;
    std::map<std::string, double> attrs;
    std::string error;
};
template <DataType> struct Native;
template <> struct Native<DataType::F32> { using type = float; };
template <> struct Native<DataType::F64> { using type = double; };
template <> struct Native<DataType::S32> { using type = int32_t; };
template <> struct Native<DataType::C64> { struct type { float re, im; }; };
}  // namespace stub

enum class ErrorCode : uint8_t { kOk = 0, kInvalidArgument = 3, kInternal = 13 };

class Error {
 public:
    Error() = default;
    Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
    static Error Success() { return Error(); }
    bool success() const { return code_ == ErrorCode::kOk; }
    bool failure() const { return !success(); }
    const std::string& message() const { return message_; }
 private:
    ErrorCode code_ = ErrorCode::kOk;
    std::string message_;
};

template <DataType dtype>
class Buffer {
 public:
    using T = typename stub::Native<dtype>::type;
    Buffer() = default;
    explicit Buffer(const stub::RawBuffer* b) : b_(b) {}
    void* untyped_data() const { return b_->data; }
    T* typed_data() const { return static_cast<T*>(b_->data); }
    size_t element_count() const { size_t n = 1; for (int64_t d : b_->dims) n *= (size_t)d; return n; }
    size_t size_bytes() const { return element_count() * sizeof(T); }
    const std::vector<int64_t>& dimensions() const { return b_->dims; }
 private:
    const stub::RawBuffer* b_ = nullptr;
};

template <typename T>
class Result {
 public:
    Result() = default;
    explicit Result(T v) : v_(v) {}
    T& operator*() { return v_; }
    T* operator->() { return &v_; }
 private:
    T v_;
};
template <DataType dtype> using ResultBuffer = Result<Buffer<dtype>>;

template <typename T> struct PlatformStream {};

namespace stub {
template <typename T> struct CtxTag {};
template <typename T> struct ArgTag {};
template <typename T> struct RetTag {};
template <typename T> struct AttrTag {};

struct DecodeState { size_t arg = 0, ret = 0, attr = 0; const std::vector<std::string>* names; bool ok = true; };

template <typename Tag> struct Decode;
template <typename S> struct Decode<CtxTag<PlatformStream<S>>> {
    static S get(CallFrame* f, DecodeState&) { return reinterpret_cast<S>(f->stream); }
};
template <DataType d> struct Decode<ArgTag<Buffer<d>>> {
    static Buffer<d> get(CallFrame* f, DecodeState& s) {
        if (s.arg >= f->args.size() || f->args[s.arg].dtype != (int)d) { s.ok = false; f->error = "argument " + std::to_string(s.arg) + ": wrong count or dtype"; static RawBuffer z{nullptr, 0, {0}}; ++s.arg; return Buffer<d>(&z); }
        return Buffer<d>(&f->args[s.arg++]);
    }
};
template <DataType d> struct Decode<RetTag<Buffer<d>>> {
    static Result<Buffer<d>> get(CallFrame* f, DecodeState& s) {
        if (s.ret >= f->rets.size() || f->rets[s.ret].dtype != (int)d) { s.ok = false; f->error = "result " + std::to_string(s.ret) + ": wrong count or dtype"; static RawBuffer z{nullptr, 0, {0}}; ++s.ret; return Result<Buffer<d>>(Buffer<d>(&z)); }
        return Result<Buffer<d>>(Buffer<d>(&f->rets[s.ret++]));
    }
};
template <> struct Decode<AttrTag<double>> {
    static double get(CallFrame* f, DecodeState& s) {
        const std::string& n = (*s.names)[s.attr++];
        auto it = f->attrs.find(n);
        if (it == f->attrs.end()) { s.ok = false; f->error = "missing attribute " + n; return 0.0; }
        return it->second;
    }
};

template <typename Fn, typename... Tags>
struct Handler {
    Fn fn; std::vector<std::string> names;
    int Call(CallFrame* f) const {
        DecodeState s; s.names = &names;
        // braced initialisation: the decoders run left to right
        std::tuple<decltype(Decode<Tags>::get(f, s))...> vals{Decode<Tags>::get(f, s)...};
        if (!s.ok) return 1;
        Error e = std::apply(fn, std::move(vals));
        if (e.failure()) { f->error = e.message(); return 1; }
        return 0;
    }
};

template <typename... Tags>
struct Binding {
    std::vector<std::string> names;
    template <typename T> Binding<Tags..., CtxTag<T>> Ctx() && { return {std::move(names)}; }
    template <typename T> Binding<Tags..., ArgTag<T>> Arg() && { return {std::move(names)}; }
    template <typename T> Binding<Tags..., RetTag<T>> Ret() && { return {std::move(names)}; }
    template <typename T> Binding<Tags..., AttrTag<T>> Attr(std::string name) && { names.push_back(std::move(name)); return {std::move(names)}; }
    template <typename Fn> Handler<Fn, Tags...> To(Fn fn) && { return {fn, std::move(names)}; }
};
}  // namespace stub

struct Ffi {
    static stub::Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

// same spelling as the real macro; the symbol takes the test double's call frame instead of XLA_FFI_CallFrame
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(fn, impl, binding)                       \
    extern "C" int fn(::xla::ffi::stub::CallFrame* frame) {                    \
        static auto* handler = new auto((binding).To(impl));                   \
        return handler->Call(frame);                                           \
    }
