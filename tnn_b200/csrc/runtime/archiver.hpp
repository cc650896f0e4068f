// Binary archiver: Sizer / Writer / Reader over a contiguous byte buffer.
//
// Encoding rules (wire + checkpoint compatible with the reference's archiver,
// include/common/archiver.hpp:51-152, include/distributed/io.hpp:19-86):
//   primitives / enums : raw host-endian
//   std::string        : u64 length + bytes
//   vector<size_t>     : u64 count + raw u64[]
//   Tensor             : u32 dtype tag, shape vector, raw size*sizeof(dtype) bytes
//   Variant            : u32 alternative index + value
//   PacketHeader       : version, type, endianness, packet_length, msg_length, msg_serial_id,
//                        packet_offset, total_packets, compression  -- read in the SAME order
//                        (the reference reads version,endianness first; that asymmetry is fixed).
// A Reader constructed with `swap=true` byte-swaps every multi-byte primitive (endianness
// mismatch between peers).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace tnn {

inline bool host_is_little_endian() {
  const uint16_t x = 1;
  return *reinterpret_cast<const uint8_t*>(&x) == 1;
}

template <typename T>
inline T bswap_value(T v) {
  uint8_t* p = reinterpret_cast<uint8_t*>(&v);
  for (size_t i = 0; i < sizeof(T) / 2; ++i) std::swap(p[i], p[sizeof(T) - 1 - i]);
  return v;
}

class Sizer {
 public:
  size_t size() const { return n_; }
  template <typename T>
  void prim(const T&) { n_ += sizeof(T); }
  void bytes(const void*, size_t n) { n_ += n; }
  void str(const std::string& s) { n_ += 8 + s.size(); }
  void sizes(const std::vector<uint64_t>& v) { n_ += 8 + 8 * v.size(); }
  void tensor(uint32_t, const std::vector<uint64_t>& shape, size_t nbytes) {
    n_ += 4;
    sizes(shape);
    n_ += nbytes;
  }

 private:
  size_t n_ = 0;
};

class Writer {
 public:
  explicit Writer(size_t reserve = 0) { buf_.reserve(reserve); }
  const std::vector<uint8_t>& buffer() const { return buf_; }
  std::vector<uint8_t>& buffer() { return buf_; }
  size_t size() const { return buf_.size(); }
  template <typename T>
  void prim(const T& v) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(&v);
    buf_.insert(buf_.end(), p, p + sizeof(T));
  }
  void bytes(const void* p, size_t n) {
    const uint8_t* b = reinterpret_cast<const uint8_t*>(p);
    buf_.insert(buf_.end(), b, b + n);
  }
  void str(const std::string& s) {
    prim<uint64_t>(s.size());
    bytes(s.data(), s.size());
  }
  void sizes(const std::vector<uint64_t>& v) {
    prim<uint64_t>(v.size());
    bytes(v.data(), v.size() * 8);
  }
  void tensor(uint32_t dtype, const std::vector<uint64_t>& shape, const void* data, size_t nbytes) {
    prim<uint32_t>(dtype);
    sizes(shape);
    bytes(data, nbytes);
  }

 private:
  std::vector<uint8_t> buf_;
};

class Reader {
 public:
  Reader(const uint8_t* data, size_t n, bool swap = false) : p_(data), n_(n), swap_(swap) {}
  size_t offset() const { return off_; }
  size_t remaining() const { return n_ - off_; }
  template <typename T>
  T prim() {
    need(sizeof(T));
    T v;
    std::memcpy(&v, p_ + off_, sizeof(T));
    off_ += sizeof(T);
    return (swap_ && sizeof(T) > 1) ? bswap_value(v) : v;
  }
  const uint8_t* bytes(size_t n) {
    need(n);
    const uint8_t* r = p_ + off_;
    off_ += n;
    return r;
  }
  std::string str() {
    uint64_t n = prim<uint64_t>();
    const uint8_t* b = bytes(n);
    return std::string(reinterpret_cast<const char*>(b), n);
  }
  std::vector<uint64_t> sizes() {
    uint64_t n = prim<uint64_t>();
    std::vector<uint64_t> v(n);
    for (uint64_t i = 0; i < n; ++i) v[i] = prim<uint64_t>();
    return v;
  }
  bool swap() const { return swap_; }

 private:
  void need(size_t n) const {
    if (n > n_ - off_) throw std::out_of_range("tnn::Reader: read past end of buffer");   // off_ <= n_ always; no wrap
  }
  const uint8_t* p_;
  size_t n_;
  size_t off_ = 0;
  bool swap_;
};

// Packet framing for the control-plane transport (reference: include/distributed/packet.hpp:10-70)
struct PacketHeader {
  uint8_t version = 1;
  uint8_t type = 0;
  uint8_t endianness = 0;  // 1 = little
  uint64_t packet_length = 0;
  uint64_t msg_length = 0;
  uint64_t msg_serial_id = 0;
  uint64_t packet_offset = 0;
  uint32_t total_packets = 1;
  uint8_t compression = 0;  // NONE
  static constexpr size_t kBytes = 3 + 8 * 4 + 4 + 1;
  void write(Writer& w) const {
    w.prim(version); w.prim(type); w.prim(endianness); w.prim(packet_length); w.prim(msg_length);
    w.prim(msg_serial_id); w.prim(packet_offset); w.prim(total_packets); w.prim(compression);
  }
  static PacketHeader read(const uint8_t* p, size_t n) {
    PacketHeader h;
    if (n < kBytes) throw std::out_of_range("PacketHeader: short buffer");
    h.version = p[0];
    h.type = p[1];
    h.endianness = p[2];
    const bool swap = (h.endianness != 0) != host_is_little_endian();
    Reader r(p + 3, n - 3, swap);
    h.packet_length = r.prim<uint64_t>();
    h.msg_length = r.prim<uint64_t>();
    h.msg_serial_id = r.prim<uint64_t>();
    h.packet_offset = r.prim<uint64_t>();
    h.total_packets = r.prim<uint32_t>();
    h.compression = r.prim<uint8_t>();
    return h;
  }
};

}  // namespace tnn
