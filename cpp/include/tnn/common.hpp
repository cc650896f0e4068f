// Small value types of the reference's `common` / `type` / `distributed` layers that other headers build on
// (reference: include/common/variant.hpp:7-159, blob.hpp:10-22, endian.hpp:6-32, include/type/type.hpp:226-286 DISPATCH_*,
// include/distributed/load_tracker.hpp:10-21, include/nn/train.hpp:31 ProfilerType, include/distributed/packet.hpp:10).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <variant>

#include "runtime/archiver.hpp"
#include "tnn/device.hpp"
#include "tnn/json.hpp"
#include "tnn/type.hpp"

namespace tnn {

// ---- endianness ------------------------------------------------------------------------------------------------------
enum class Endianness : uint8_t { BIG = 0, LITTLE = 1 };
inline Endianness host_endianness() { return host_is_little_endian() ? Endianness::LITTLE : Endianness::BIG; }
template <class T>
inline T bswap(T v) {
  static_assert(std::is_trivially_copyable<T>::value, "bswap: trivially copyable types only");
  return bswap_value(v);
}
// in-place swap of `count` elements of `elem_size` bytes (ops::bswap of the reference, host side)
inline void bswap_buffer(void* data, size_t count, size_t elem_size) {
  if (elem_size < 2) return;
  auto* p = static_cast<uint8_t*>(data);
  for (size_t i = 0; i < count; ++i, p += elem_size)
    for (size_t a = 0, b = elem_size - 1; a < b; ++a, --b) std::swap(p[a], p[b]);
}

// ---- Blob: a typed, non-owning range on a device ------------------------------------------------------------------
template <class T>
struct Blob {
  using value_type = T;
  T* ptr;
  uint64_t count;
  Device device;
  uint64_t bytes() const { return count * sizeof(T); }
};
template <class T, class = std::enable_if_t<!std::is_void<T>::value>>
Blob<T> make_blob(T* data, uint64_t count, const Device& device = getHost()) {
  return Blob<T>{data, count, device};
}

// ---- Variant: the tagged union the wire format is written in terms of (u32 index + value) ----------------------------
// std::variant does the storage; the reference's interface (index_of<T>(), is<T>(), get<T>(), visit, index()) is kept so that
// payload code reads the same, and archive() writes the reference's encoding through any archiver with prim() / value().
template <class... Types>
class Variant {
 public:
  Variant() = default;
  template <class T, class = std::enable_if_t<(std::is_same<std::decay_t<T>, Types>::value || ...)>>
  Variant(T&& v) : v_(std::forward<T>(v)) {}          // NOLINT: implicit by design, like the reference
  template <class T>
  static constexpr size_t index_of() {
    static_assert((std::is_same<T, Types>::value || ...), "Type not found in Variant");
    size_t index = 0, match = 0;
    ((std::is_same<T, Types>::value ? (match = index) : ++index), ...);
    return match;
  }
  static constexpr size_t size() { return sizeof...(Types); }
  size_t index() const { return v_.index(); }
  template <class T> bool is() const { return std::holds_alternative<T>(v_); }
  template <class T> T& get() {
    if (!is<T>()) throw std::runtime_error("Variant: holds alternative " + std::to_string(index()) + ", not " + std::to_string(index_of<T>()));
    return std::get<T>(v_);
  }
  template <class T> const T& get() const { return const_cast<Variant*>(this)->get<T>(); }
  template <class T> T* get_if() { return std::get_if<T>(&v_); }
  template <class T, class... A> T& emplace(A&&... a) { return v_.template emplace<T>(std::forward<A>(a)...); }
  template <class F> decltype(auto) visit(F&& f) { return std::visit(std::forward<F>(f), v_); }
  template <class F> decltype(auto) visit(F&& f) const { return std::visit(std::forward<F>(f), v_); }
  // construct alternative `idx` by default (the read side of the wire format: index first, then the value)
  void set_index(size_t idx) {
    if (idx >= sizeof...(Types)) throw std::out_of_range("Variant: index " + std::to_string(idx) + " of " + std::to_string(sizeof...(Types)));
    set_index_impl(idx, std::index_sequence_for<Types...>{});
  }
  bool operator==(const Variant& o) const { return v_ == o.v_; }

 private:
  template <size_t... I>
  void set_index_impl(size_t idx, std::index_sequence<I...>) {
    ((idx == I ? (void)v_.template emplace<I>() : (void)0), ...);
  }
  std::variant<Types...> v_;
};

// ---- dtype dispatch ---------------------------------------------------------------------------------------------------
// dispatch_float_dtype(dt, f) calls f(T{}) with the host carrier type of a floating dtype (float, double, bf16_t, fp16_t);
// dispatch_any_dtype adds the integer / byte types.  TNN_DISPATCH_DTYPE(dt, T, body) is the macro spelling of the same.
template <class F>
decltype(auto) dispatch_float_dtype(DType_t dt, F&& f) {
  switch (dt) {
    case DType_t::FP32: return f(float{});
    case DType_t::FP64: return f(double{});
    case DType_t::BF16: return f(bf16_t{});
    case DType_t::FP16: return f(fp16_t{});
    default: throw std::runtime_error(std::string("dispatch_float_dtype: ") + dtype_to_string(dt) + " is not a floating type");
  }
}
template <class F>
void dispatch_any_dtype(DType_t dt, F&& f) {
  switch (dt) {
    case DType_t::FP32: f(float{}); return;
    case DType_t::FP64: f(double{}); return;
    case DType_t::BF16: f(bf16_t{}); return;
    case DType_t::FP16: f(fp16_t{}); return;
    case DType_t::INT32_T: f(int32_t{}); return;
    case DType_t::INT64_T: f(int64_t{}); return;
    case DType_t::SIZE_T: f(uint64_t{}); return;
    case DType_t::UINT8_T: case DType_t::BYTE: f(uint8_t{}); return;
    case DType_t::BOOL: f(bool{}); return;
    default: throw std::runtime_error("dispatch_any_dtype: unknown dtype");
  }
}
#define TNN_DISPATCH_DTYPE(dt, T, ...) ::tnn::dispatch_float_dtype((dt), [&](auto tnn_tag__) { using T = decltype(tnn_tag__); __VA_ARGS__; })
#define TNN_DISPATCH_ANY_DTYPE(dt, T, ...) ::tnn::dispatch_any_dtype((dt), [&](auto tnn_tag__) { using T = decltype(tnn_tag__); __VA_ARGS__; })

// ---- LoadTracker: what a stage reports about itself (REPORT_LOAD -> LOAD_REPORT) ---------------------------------------
struct LoadTracker {
  float avg_forward_time_ = 0;       // ms per forward job
  float avg_backward_time_ = 0;      // ms per backward job
  float avg_cpu_utilization_ = -1;   // CPU seconds / wall seconds of the worker process since start, -1 when unavailable
  float max_memory_usage_ = -1;      // peak resident set in MB, -1 when unavailable
  float max_device_memory_usage_ = 0;  // MB in use on the stage's GPU at report time (0 on the CPU)
  Json to_json() const {
    Json j = Json::object();
    j["avg_forward_time"] = (double)avg_forward_time_;
    j["avg_backward_time"] = (double)avg_backward_time_;
    j["avg_cpu_utilization"] = (double)avg_cpu_utilization_;
    j["max_memory_usage"] = (double)max_memory_usage_;
    j["max_device_memory_usage"] = (double)max_device_memory_usage_;
    return j;
  }
  static LoadTracker from_json(const Json& j) {
    LoadTracker t;
    t.avg_forward_time_ = (float)j.value("avg_forward_time", 0.0);
    t.avg_backward_time_ = (float)j.value("avg_backward_time", 0.0);
    t.avg_cpu_utilization_ = (float)j.value("avg_cpu_utilization", -1.0);
    t.max_memory_usage_ = (float)j.value("max_memory_usage", -1.0);
    t.max_device_memory_usage_ = (float)j.value("max_device_memory_usage", 0.0);
    return t;
  }
};

// ---- enums of TrainingConfig / PacketHeader ------------------------------------------------------------------------------
enum class ProfilerType { NONE = 0, NORMAL = 1, CUMULATIVE = 2 };
inline const char* profiler_type_to_string(ProfilerType t) { return t == ProfilerType::NORMAL ? "NORMAL" : t == ProfilerType::CUMULATIVE ? "CUMULATIVE" : "NONE"; }
inline ProfilerType string_to_profiler_type(std::string s) {
  for (auto& c : s) c = (char)toupper((unsigned char)c);
  if (s == "NORMAL") return ProfilerType::NORMAL;
  if (s == "CUMULATIVE") return ProfilerType::CUMULATIVE;
  if (s.empty() || s == "NONE") return ProfilerType::NONE;
  throw std::runtime_error("unknown profiler type '" + s + "' (NONE | NORMAL | CUMULATIVE)");
}
// PacketHeader::compression: declared by the reference (packet.hpp:10), never implemented there; frames here are always NONE
enum class CompressionType : uint8_t { NONE = 0, ZSTD = 1, LZ4 = 2 };

}  // namespace tnn
