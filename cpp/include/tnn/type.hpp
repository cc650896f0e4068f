// Runtime dtype enum, sizes, strings and per-dtype epsilon (reference: include/type/type.hpp:76-224,
// src/type/type.cpp:8-11).  Device arithmetic types are CUDA's __nv_bfloat16 / __half; the host carries bf16 /
// fp16 values as raw 16-bit words with the conversion helpers below (round to nearest even).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

namespace tnn {

enum class DType_t : uint32_t { BYTE = 0, UINT8_T, BOOL, FP16, BF16, FP32, FP64, INT32_T, INT64_T, SIZE_T, UNKNOWN };

inline size_t get_dtype_size(DType_t t) {
  switch (t) {
    case DType_t::BYTE: case DType_t::UINT8_T: case DType_t::BOOL: return 1;
    case DType_t::FP16: case DType_t::BF16: return 2;
    case DType_t::FP32: case DType_t::INT32_T: return 4;
    case DType_t::FP64: case DType_t::INT64_T: case DType_t::SIZE_T: return 8;
    default: throw std::runtime_error("get_dtype_size: unknown dtype");
  }
}
inline const char* dtype_to_string(DType_t t) {
  switch (t) {
    case DType_t::BYTE: return "BYTE"; case DType_t::UINT8_T: return "UINT8"; case DType_t::BOOL: return "BOOL";
    case DType_t::FP16: return "FP16"; case DType_t::BF16: return "BF16"; case DType_t::FP32: return "FP32";
    case DType_t::FP64: return "FP64"; case DType_t::INT32_T: return "INT32"; case DType_t::INT64_T: return "INT64";
    case DType_t::SIZE_T: return "SIZE_T"; default: return "UNKNOWN";
  }
}
inline DType_t string_to_dtype(std::string s) {
  for (auto& c : s) c = (char)toupper((unsigned char)c);
  if (s == "FP32" || s == "FLOAT" || s == "FLOAT32") return DType_t::FP32;
  if (s == "BF16" || s == "BFLOAT16") return DType_t::BF16;
  if (s == "FP16" || s == "HALF" || s == "FLOAT16") return DType_t::FP16;
  if (s == "FP64" || s == "DOUBLE" || s == "FLOAT64") return DType_t::FP64;
  if (s == "INT32" || s == "INT") return DType_t::INT32_T;
  if (s == "INT64" || s == "LONG") return DType_t::INT64_T;
  if (s == "UINT8") return DType_t::UINT8_T;
  if (s == "BYTE") return DType_t::BYTE;
  if (s == "BOOL") return DType_t::BOOL;
  if (s == "SIZE_T") return DType_t::SIZE_T;
  throw std::runtime_error("string_to_dtype: unknown dtype '" + s + "'");
}
// src/type/type.cpp:8-11
inline double dtype_eps(DType_t t) {
  switch (t) {
    case DType_t::FP16: return 1e-3;
    case DType_t::BF16: return 1e-2;
    case DType_t::FP32: return 1e-5;
    case DType_t::FP64: return 1e-8;
    default: return 0.0;
  }
}
template <class T> struct dtype_of_t;
template <> struct dtype_of_t<float> { static constexpr DType_t value = DType_t::FP32; };
template <> struct dtype_of_t<double> { static constexpr DType_t value = DType_t::FP64; };
template <> struct dtype_of_t<int32_t> { static constexpr DType_t value = DType_t::INT32_T; };
template <> struct dtype_of_t<int64_t> { static constexpr DType_t value = DType_t::INT64_T; };
template <> struct dtype_of_t<uint8_t> { static constexpr DType_t value = DType_t::UINT8_T; };
template <> struct dtype_of_t<bool> { static constexpr DType_t value = DType_t::BOOL; };
template <class T> constexpr DType_t dtype_of() { return dtype_of_t<T>::value; }

// ---- host-side 16-bit float carriers -------------------------------------------------------------
struct bf16_t { uint16_t bits; };
struct fp16_t { uint16_t bits; };
template <> struct dtype_of_t<bf16_t> { static constexpr DType_t value = DType_t::BF16; };
template <> struct dtype_of_t<fp16_t> { static constexpr DType_t value = DType_t::FP16; };

inline bf16_t float_to_bf16(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x7FFFFFu)) return bf16_t{(uint16_t)((u >> 16) | 0x40)};  // quiet NaN
  u += 0x7FFFu + ((u >> 16) & 1u);                                                                          // RNE
  return bf16_t{(uint16_t)(u >> 16)};
}
inline float bf16_to_float(bf16_t h) {
  const uint32_t u = (uint32_t)h.bits << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
inline fp16_t float_to_fp16(float f) {
  uint32_t x;
  std::memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  const int32_t exp = (int32_t)((x >> 23) & 0xFF) - 127 + 15;
  uint32_t man = x & 0x7FFFFFu;
  if (((x >> 23) & 0xFF) == 0xFF) return fp16_t{(uint16_t)(sign | 0x7C00u | (man ? 0x200u : 0))};
  if (exp >= 31) return fp16_t{(uint16_t)(sign | 0x7C00u)};
  if (exp <= 0) {
    if (exp < -10) return fp16_t{(uint16_t)sign};
    man |= 0x800000u;
    const int shift = 14 - exp;
    uint32_t h = man >> shift;
    const uint32_t rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1))) ++h;
    return fp16_t{(uint16_t)(sign | h)};
  }
  uint32_t h = ((uint32_t)exp << 10) | (man >> 13);
  const uint32_t rem = man & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;
  return fp16_t{(uint16_t)(sign | h)};
}
inline float fp16_to_float(fp16_t h) {
  const uint32_t sign = ((uint32_t)h.bits & 0x8000u) << 16;
  uint32_t exp = (h.bits >> 10) & 0x1F, man = h.bits & 0x3FFu, u;
  if (exp == 0) {
    if (man == 0) u = sign;
    else {
      int e = -1;
      do { ++e; man <<= 1; } while (!(man & 0x400u));
      u = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
    }
  } else if (exp == 31) u = sign | 0x7F800000u | (man << 13);
  else u = sign | ((exp - 15 + 127) << 23) | (man << 13);
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

}  // namespace tnn
