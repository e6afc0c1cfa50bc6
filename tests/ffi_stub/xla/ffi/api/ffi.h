// Stand-in for XLA's "xla/ffi/api/ffi.h" (not in this image): just enough of the public surface that
// mccube_b200/csrc/xla_ffi_shim.cc uses, so that tests/test_ffi_shim.py can COMPILE the shim on a CPU box and
// have every C-ABI call type-checked against include/mccube_b200.h, and every binding checked against the
// signature of its handler (static_assert below).  Test infrastructure only -- nothing here runs.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

namespace xla::ffi {

enum DataType { U8, U32, S32, U64, F32, F64 };
template <DataType> struct NativeOf;
template <> struct NativeOf<U8> { using type = uint8_t; };
template <> struct NativeOf<U32> { using type = uint32_t; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<U64> { using type = uint64_t; };
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<F64> { using type = double; };

template <typename T>
class Span {
 public:
  const T* begin() const { return data_; }
  const T* end() const { return data_ + size_; }
  size_t size() const { return size_; }
  const T& operator[](size_t i) const { return data_[i]; }
 private:
  const T* data_ = nullptr;
  size_t size_ = 0;
};

template <DataType dtype>
class Buffer {
 public:
  using T = typename NativeOf<dtype>::type;
  T* typed_data() const { return data_; }
  Span<const int64_t> dimensions() const { return dims_; }
  size_t element_count() const { return count_; }
  size_t size_bytes() const { return count_ * sizeof(T); }
 private:
  T* data_ = nullptr;
  Span<const int64_t> dims_;
  size_t count_ = 0;
};

template <typename B>
class Result {
 public:
  B* operator->() { return &b_; }
  B& operator*() { return b_; }
 private:
  B b_;
};

enum class ErrorCode { kOk, kInvalidArgument, kInternal };
class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string msg) : code_(code), msg_(std::move(msg)) {}
  static Error Success() { return Error(); }
 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

template <typename T> struct PlatformStream {};
enum class Traits { kCmdBufferCompatible };

// The binding only records the C++ parameter list the decoded call frame would be passed as.
template <typename... Ps>
struct Binding {
  template <typename C> auto Ctx() const { return CtxOf<C>(); }
  template <typename B> Binding<Ps..., B> Arg() const { return {}; }
  template <typename T> Binding<Ps..., T> Attr(const char*) const { return {}; }
  template <typename B> Binding<Ps..., Result<B>> Ret() const { return {}; }
  template <typename F> static constexpr bool Matches() { return std::is_invocable_r_v<Error, F, Ps...>; }
 private:
  template <typename T> static Binding<Ps..., T> CtxOf(PlatformStream<T>*) { return {}; }
  template <typename C> static auto CtxOf() { return CtxOf(static_cast<C*>(nullptr)); }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding, ...)                                              \
  static_assert(decltype(binding)::template Matches<decltype(&impl)>(), #sym ": binding and handler differ"); \
  extern "C" const void* sym() { return reinterpret_cast<const void*>(&impl); }
