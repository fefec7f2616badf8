// TEST-ONLY stub of the part of the public XLA typed-FFI C++ API (xla/ffi/api/ffi.h) that csrc/xla_ffi_shim.cc uses,
// written from the API's documentation because neither jax nor the XLA headers are installable in this image.  It lets
// `pytest -m "not gpu"` COMPILE the shim (names, arity and types of every handler against include/exenv_b200.h); it
// proves nothing about the real headers and is never part of the product.
#pragma once
#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <vector>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla {
namespace ffi {

enum class ErrorCode { kOk, kInvalidArgument, kInternal };

class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

template <typename T>
struct Span {
  const T *p = nullptr;
  size_t n = 0;
  const T &operator[](size_t i) const { return p[i]; }
  size_t size() const { return n; }
};

class AnyBuffer {
 public:
  void *untyped_data() const { return data_; }
  Span<int64_t> dimensions() const { return Span<int64_t>{dims_.data(), dims_.size()}; }
  size_t element_count() const {
    size_t n = 1;
    for (int64_t d : dims_) n *= (size_t)d;
    return n;
  }
  void *data_ = nullptr;
  std::vector<int64_t> dims_;
};

template <typename T>
class Result {
 public:
  T *operator->() { return &v_; }
  T &operator*() { return v_; }
  T v_;
};

class RemainingArgs {
 public:
  size_t size() const { return v_.size(); }
  template <typename T>
  std::optional<T> get(size_t i) const {
    if (i >= v_.size()) return std::nullopt;
    return v_[i];
  }
  std::vector<AnyBuffer> v_;
};

class RemainingRets {
 public:
  size_t size() const { return v_.size(); }
  template <typename T>
  std::optional<Result<T>> get(size_t i) const {
    if (i >= v_.size()) return std::nullopt;
    return Result<T>{v_[i]};
  }
  std::vector<AnyBuffer> v_;
};

template <typename T>
struct PlatformStream {};

// Binding builder: records the handler signature as a type list so that the handler function is type-checked.
template <typename... Ts>
struct Binding {
  template <typename T>
  Binding<Ts..., T> Ctx() const { return {}; }
  Binding<Ts..., ::xla::ffi::RemainingArgs> RemainingArgs() const { return {}; }
  Binding<Ts..., ::xla::ffi::RemainingRets> RemainingRets() const { return {}; }
  template <typename T>
  Binding<Ts..., T> Attr(const char *) const { return {}; }
};
template <typename T>
struct CtxArg { using type = T; };
template <typename T>
struct CtxArg<PlatformStream<T>> { using type = T; };

struct Ffi {
  static Binding<> Bind() { return {}; }
};

template <typename... Ts>
constexpr bool check_handler(Error (*)(typename CtxArg<Ts>::type...), Binding<Ts...>) { return true; }

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                   \
  static_assert(sizeof(decltype(::xla::ffi::check_handler(impl, binding))) > 0, "signature"); \
  extern "C" XLA_FFI_Error *name(XLA_FFI_CallFrame *) { return nullptr; }
