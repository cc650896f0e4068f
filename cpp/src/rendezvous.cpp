// RoCECommunicator: the one-sided rendezvous transport of the C++ host (see tnn/distributed.hpp).  Frames are those of
// tnn_b200/distributed/rendezvous.py: {u8 kind, u64 a, u64 b} little endian + tail.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "runtime/symmetric_memory.hpp"
#include "tnn/distributed.hpp"

namespace tnn {

namespace {
enum : uint8_t { K_INLINE = 0, K_PREPARE = 1, K_READY = 2, K_DONE = 3, K_CANCEL = 4, K_GPREPARE = 5, K_GREADY = 6, K_GDONE = 7 };
constexpr uint64_t kNoBlock = OffsetFreeList::kNone;
constexpr size_t kHdr = 17;

std::string host_token() {
  std::ifstream f("/proc/sys/kernel/random/boot_id");
  std::string s;
  if (f && std::getline(f, s)) {
    while (!s.empty() && isspace((unsigned char)s.back())) s.pop_back();
    return s;
  }
  char buf[256] = {};
  gethostname(buf, sizeof(buf) - 1);
  return buf;
}
std::vector<uint8_t> frame(uint8_t kind, uint64_t a, uint64_t b, const void* tail = nullptr, size_t n = 0) {
  std::vector<uint8_t> f(kHdr + n);
  f[0] = kind;
  std::memcpy(&f[1], &a, 8);
  std::memcpy(&f[9], &b, 8);
  if (n) std::memcpy(&f[kHdr], tail, n);
  return f;
}
std::string to_hex(const void* p, size_t n) {
  static const char* d = "0123456789abcdef";
  std::string s(2 * n, '0');
  const uint8_t* b = static_cast<const uint8_t*>(p);
  for (size_t i = 0; i < n; ++i) {
    s[2 * i] = d[b[i] >> 4];
    s[2 * i + 1] = d[b[i] & 15];
  }
  return s;
}
bool from_hex(const std::string& s, void* out, size_t n) {
  if (s.size() != 2 * n) return false;
  auto v = [](char c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1; };
  for (size_t i = 0; i < n; ++i) {
    const int hi = v(s[2 * i]), lo = v(s[2 * i + 1]);
    if (hi < 0 || lo < 0) return false;
    static_cast<uint8_t*>(out)[i] = (uint8_t)(hi * 16 + lo);
  }
  return true;
}
std::vector<std::string> split(const std::string& s, char sep) {
  std::vector<std::string> out;
  std::string cur;
  for (char c : s) {
    if (c == sep) {
      out.push_back(cur);
      cur.clear();
    } else {
      cur.push_back(c);
    }
  }
  out.push_back(cur);
  return out;
}
#define RDV_CUDA(call)                                                                                                \
  do {                                                                                                                \
    cudaError_t e_ = (call);                                                                                          \
    if (e_ != cudaSuccess) throw DistributedError(std::string("RoCECommunicator: " #call ": ") + cudaGetErrorString(e_)); \
  } while (0)
}  // namespace

namespace {
// a POSIX shared-memory segment under /dev/shm (the creator unlinks it)
struct ShmSegment {
  std::string name;
  uint8_t* base = nullptr;
  size_t bytes = 0;
  bool created = false;
  bool open(const std::string& n, size_t size, bool create) {
    name = n;
    created = create;
    const std::string path = "/dev/shm/" + n;
    const int fd = ::open(path.c_str(), create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return false;
    if (create && ftruncate(fd, (off_t)size) != 0) {
      ::close(fd);
      return false;
    }
    struct stat st;
    if (fstat(fd, &st) != 0) {
      ::close(fd);
      return false;
    }
    bytes = (size_t)st.st_size;
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    ::close(fd);
    if (p == MAP_FAILED) return false;
    base = static_cast<uint8_t*>(p);
    return true;
  }
  void close() {
    if (base) munmap(base, bytes);
    base = nullptr;
    if (created && !name.empty()) unlink(("/dev/shm/" + name).c_str());
    created = false;
  }
};
}  // namespace

struct RoCECommunicator::Impl {
  size_t slab_bytes = 0, inline_bytes = 0, host_slab_bytes = 64u << 20;
  std::string token = host_token();
  // host slab: a shared-memory segment peers on this node write CPU messages into (PREPARE / READY_TO_WRITE / write / DONE)
  ShmSegment shm;
  OffsetFreeList hslab{0, 256};              // best fit + coalescing (the IbvAllocator role: one registered range)
  std::map<std::string, ShmSegment> peer_shm;               // segment name -> mapping
  std::map<uint64_t, std::pair<std::vector<uint8_t>, std::string>> pending_host;   // id -> (serialised message, peer key)
  std::vector<std::tuple<std::string, uint64_t, uint64_t>> waiting;                // PREPAREs that did not fit yet
  // device slab (OffsetFreeList, 512-byte blocks)
  int dev = -1;
  uint8_t* dptr = nullptr;
  cudaIpcMemHandle_t handle{};
  std::mutex mu;
  OffsetFreeList dslab{0, 512};
  std::vector<std::pair<cudaEvent_t, uint64_t>> released;  // blocks waiting for the consumer stream
  std::map<std::string, void*> opened;                     // peer ipc handle (hex) -> mapped base
  cudaStream_t copy_stream = nullptr;
  struct Pending {
    Message msg;
    Endpoint to;
  };
  std::map<uint64_t, Pending> pending;
  uint64_t next_id = 1;
  Stats st;

  uint64_t alloc(uint64_t n) {
    const uint64_t off = dslab.allocate(n);
    st.slab_peak = std::max<uint64_t>(st.slab_peak, dslab.peak());
    return off;
  }
  void release(uint64_t off) { dslab.release(off); }
  void reap() {   // mu held
    for (size_t i = 0; i < released.size();) {
      if (released[i].first == nullptr || cudaEventQuery(released[i].first) == cudaSuccess) {
        if (released[i].first) cudaEventDestroy(released[i].first);
        release(released[i].second);
        released[i] = released.back();
        released.pop_back();
      } else {
        ++i;
      }
    }
  }
};

RoCECommunicator::RoCECommunicator(const std::string& host, int port, int spp, const Device* land_on, const std::string& name,
                                   size_t device_slab_bytes, size_t inline_bytes, size_t host_slab_bytes)
    : TCPCommunicator(host, port, spp, land_on, name), impl_(std::make_shared<Impl>()) {
  impl_->slab_bytes = device_slab_bytes;
  impl_->inline_bytes = inline_bytes;
  impl_->host_slab_bytes = host_slab_bytes;
}
RoCECommunicator::~RoCECommunicator() { stop(); }
RoCECommunicator::Stats RoCECommunicator::stats() const {
  std::lock_guard<std::mutex> g(impl_->mu);
  Stats s = impl_->st;
  s.host_slab_peak = impl_->hslab.peak();
  return s;
}

void RoCECommunicator::start() {
  if (running_.exchange(true)) return;
  Impl& I = *impl_;
  const char* off = getenv("TNN_DEVICE_RENDEZVOUS");
  if (has_land_ && land_.is_gpu() && I.slab_bytes > 0 && !(off && off[0] == '0')) {
    I.dev = land_.id();
    land_.set_current();
    RDV_CUDA(cudaMalloc(reinterpret_cast<void**>(&I.dptr), I.slab_bytes));
    RDV_CUDA(cudaIpcGetMemHandle(&I.handle, I.dptr));
    RDV_CUDA(cudaStreamCreateWithFlags(&I.copy_stream, cudaStreamNonBlocking));
    I.dslab.reset(I.slab_bytes);
  }
  if (I.host_slab_bytes > 0) {
    static std::atomic<unsigned> serial{0};
    char nm[64];
    snprintf(nm, sizeof(nm), "tnn_%08x%04x%04x", (unsigned)getpid(), serial.fetch_add(1) & 0xffffu, (unsigned)(Profiler::now_ns() & 0xffff));
    if (I.shm.open(nm, I.host_slab_bytes, true)) {
      I.hslab.reset(I.host_slab_bytes);
    }
  }
  bus_->start();
  pump_ = std::thread([this] {
    Impl& I = *impl_;
    // replies go to the sender's listening address ("host:port" is how the bus names a peer); connect on first use
    auto reply = [this](const std::string& to, const std::vector<uint8_t>& f) {
      if (!bus_->has_peer(to)) {
        const size_t c = to.rfind(':');
        if (c == std::string::npos) throw DistributedError("RoCECommunicator: cannot reply to " + to);
        bus_->connect(to.substr(0, c), atoi(to.c_str() + c + 1));
      }
      bus_->send(to, f.data(), f.size());
    };
    while (running_) {
      std::string from;
      std::vector<uint8_t> buf;
      if (!bus_->recv(200, &from, &buf)) continue;
      if (buf.size() < kHdr) continue;
      const uint8_t kind = buf[0];
      uint64_t a, b;
      std::memcpy(&a, &buf[1], 8);
      std::memcpy(&b, &buf[9], 8);
      const uint8_t* tail = buf.data() + kHdr;
      const size_t tn = buf.size() - kHdr;
      try {
        if (kind == K_INLINE) {
          Message m = Message::deserialize(tail, tn, false, has_land_ ? &land_ : nullptr);
          m.sender = from;
          enqueue(std::move(m));
        } else if (kind == K_PREPARE) {
          // grant a block of the registered shared-memory slab (same host), queue the request while the slab is full, or
          // answer "no block" (other host / larger than the slab / no slab): the peer then resends the message inline
          const std::string tok(reinterpret_cast<const char*>(tail), tn);
          auto grant = [&](const std::string& to, uint64_t mid, uint64_t len) -> bool {
            uint64_t off = kNoBlock;
            const bool can = I.shm.base && tok == I.token && len <= I.host_slab_bytes;
            if (can) {
              std::lock_guard<std::mutex> g(I.mu);
              off = I.hslab.allocate(len);
              if (off == kNoBlock) return false;
            }
            const std::string t = can ? I.token + "|" + I.shm.name : std::string("-|");
            reply(to, frame(K_READY, mid, off, t.data(), t.size()));
            return true;
          };
          if (!grant(from, a, b)) {
            std::lock_guard<std::mutex> g(I.mu);
            I.waiting.emplace_back(from, a, b);
          }
        } else if (kind == K_READY) {
          std::pair<std::vector<uint8_t>, std::string> item;
          {
            std::lock_guard<std::mutex> g(I.mu);
            auto it = I.pending_host.find(a);
            if (it == I.pending_host.end()) continue;
            item = std::move(it->second);
            I.pending_host.erase(it);
          }
          bool ok = false;
          if (b != kNoBlock) {
            const auto parts = split(std::string(reinterpret_cast<const char*>(tail), tn), '|');
            if (parts.size() == 2 && parts[0] == I.token) {
              ShmSegment* seg = nullptr;
              {
                std::lock_guard<std::mutex> g(I.mu);
                auto it = I.peer_shm.find(parts[1]);
                if (it == I.peer_shm.end()) {
                  ShmSegment sg;
                  if (sg.open(parts[1], 0, false)) it = I.peer_shm.emplace(parts[1], sg).first;
                }
                if (it != I.peer_shm.end()) seg = &it->second;
              }
              if (seg && b + item.first.size() <= seg->bytes) {
                std::memcpy(seg->base + b, item.first.data(), item.first.size());        // the one-sided write
                const uint64_t len = item.first.size();
                reply(from, frame(K_DONE, a, b, &len, 8));                                 // ... and its immediate
                std::lock_guard<std::mutex> g(I.mu);
                ++I.st.rendezvous_sent;
                ok = true;
              }
            }
            if (!ok) reply(from, frame(K_CANCEL, a, b));
          }
          if (!ok) {
            reply(from, frame(K_INLINE, 0, item.first.size(), item.first.data(), item.first.size()));
            std::lock_guard<std::mutex> g(I.mu);
            ++I.st.fallback_inline;
          }
        } else if (kind == K_DONE || kind == K_CANCEL) {
          Message m;
          bool have = false;
          if (kind == K_DONE && I.shm.base && tn >= 8) {
            uint64_t len;
            std::memcpy(&len, tail, 8);
            if (b + len <= I.shm.bytes) {
              try {
                m = Message::deserialize(I.shm.base + b, (size_t)len, false, has_land_ ? &land_ : nullptr);   // copies out of the slab
                m.sender = from;
                have = true;
              } catch (const std::exception& e) {           // the block is released below either way
                std::cerr << "[" << name_ << "] dropped a malformed slab message from " << from << ": " << e.what() << std::endl;
              }
            }
          }
          std::vector<std::tuple<std::string, uint64_t, uint64_t>> retry;
          {
            std::lock_guard<std::mutex> g(I.mu);
            if (b != kNoBlock) I.hslab.release(b);
            if (kind == K_DONE) ++I.st.rendezvous_received;
            retry.swap(I.waiting);
          }
          if (have) enqueue(std::move(m));                  // counters and the block are settled before the consumer sees it
          for (auto& w : retry) {                          // blocks were freed: serve the PREPAREs that waited
            uint64_t off;
            {
              std::lock_guard<std::mutex> g(I.mu);
              off = I.hslab.allocate(std::get<2>(w));
              if (off == kNoBlock) {
                I.waiting.push_back(w);
                continue;
              }
            }
            const std::string t = I.token + "|" + I.shm.name;
            reply(std::get<0>(w), frame(K_READY, std::get<1>(w), off, t.data(), t.size()));
          }
        } else if (kind == K_GPREPARE) {
          const std::string ts(reinterpret_cast<const char*>(tail), tn);
          const std::string tok = split(ts, '|')[0];
          uint64_t off = kNoBlock;
          if (I.dptr && tok == I.token && b <= I.slab_bytes) {
            std::lock_guard<std::mutex> g(I.mu);
            I.reap();
            off = I.alloc(b);
          }
          std::string t;
          if (off != kNoBlock) t = std::to_string((long long)getpid()) + "|" + std::to_string((unsigned long long)(uintptr_t)I.dptr) + "|" + to_hex(&I.handle, sizeof(I.handle));
          const auto f = frame(K_GREADY, a, off, t.data(), t.size());
          reply(from, f);
        } else if (kind == K_GREADY) {
          Impl::Pending p;
          {
            std::lock_guard<std::mutex> g(I.mu);
            auto it = I.pending.find(a);
            if (it == I.pending.end()) continue;
            p = std::move(it->second);
            I.pending.erase(it);
          }
          bool ok = false;
          if (b != kNoBlock) {
            try {
              const auto parts = split(std::string(reinterpret_cast<const char*>(tail), tn), '|');
              if (parts.size() != 3) throw DistributedError("bad G_READY tail");
              uint8_t* base = nullptr;
              if (atoll(parts[0].c_str()) == (long long)getpid()) {
                base = reinterpret_cast<uint8_t*>((uintptr_t)strtoull(parts[1].c_str(), nullptr, 10));   // same process: the pointer is usable as is
              } else {
                std::lock_guard<std::mutex> g(I.mu);
                auto it = I.opened.find(parts[2]);
                if (it == I.opened.end()) {
                  cudaIpcMemHandle_t h;
                  if (!from_hex(parts[2], &h, sizeof(h))) throw DistributedError("bad ipc handle");
                  void* mapped = nullptr;
                  p.msg.job.data->device().set_current();
                  RDV_CUDA(cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
                  it = I.opened.emplace(parts[2], mapped).first;
                }
                base = static_cast<uint8_t*>(it->second);
              }
              const Tensor& t = p.msg.job.data;
              t->device().set_current();
              if (!I.copy_stream) RDV_CUDA(cudaStreamCreateWithFlags(&I.copy_stream, cudaStreamNonBlocking));
              cudaStream_t cs = I.copy_stream;
              RDV_CUDA(cudaMemcpyAsync(base + b, t->data(), t->nbytes(), cudaMemcpyDefault, cs));     // the one-sided write
              RDV_CUDA(cudaStreamSynchronize(cs));
              // metadata record of rendezvous.py: "<HQIIQ" + dims
              std::vector<uint8_t> meta(26 + 8 * t->shape().size());
              const uint16_t cmd = (uint16_t)p.msg.command;
              const uint64_t mb = p.msg.job.mb_id, nb = t->nbytes();
              const uint32_t dt = (uint32_t)t->data_type(), nd = (uint32_t)t->shape().size();
              std::memcpy(&meta[0], &cmd, 2);
              std::memcpy(&meta[2], &mb, 8);
              std::memcpy(&meta[10], &dt, 4);
              std::memcpy(&meta[14], &nd, 4);
              std::memcpy(&meta[18], &nb, 8);
              for (size_t i = 0; i < t->shape().size(); ++i) {
                const uint64_t d = t->shape()[i];
                std::memcpy(&meta[26 + 8 * i], &d, 8);
              }
              const auto f = frame(K_GDONE, a, b, meta.data(), meta.size());     // ... and its immediate
              reply(from, f);
              std::lock_guard<std::mutex> g(I.mu);
              ++I.st.device_rendezvous_sent;
              ok = true;
            } catch (const std::exception& e) {
              std::cerr << "[" << name_ << "] device rendezvous failed (" << e.what() << "), using the host path" << std::endl;
              const auto f = frame(K_GDONE, a, b);                               // empty metadata = cancelled
              reply(from, f);
            }
          }
          if (!ok) {
            {
              std::lock_guard<std::mutex> g(I.mu);
              ++I.st.device_fallback_host;
            }
            const std::vector<uint8_t> body = p.msg.serialize();
            const auto f = frame(K_INLINE, 0, body.size(), body.data(), body.size());
            reply(from, f);
          }
        } else if (kind == K_GDONE) {
          if (!I.dptr) continue;
          if (tn < 26) {
            std::lock_guard<std::mutex> g(I.mu);
            I.release(b);
            continue;
          }
          uint16_t cmd;
          uint64_t mb, nb;
          uint32_t dt, nd;
          std::memcpy(&cmd, tail, 2);
          std::memcpy(&mb, tail + 2, 8);
          std::memcpy(&dt, tail + 10, 4);
          std::memcpy(&nd, tail + 14, 4);
          std::memcpy(&nb, tail + 18, 8);
          if (dt >= (uint32_t)DType_t::UNKNOWN || nd > 16 || tn < 26 + 8ull * nd) throw DistributedError("bad G_DONE metadata");
          Shape shape;
          for (uint32_t i = 0; i < nd; ++i) {
            uint64_t d;
            std::memcpy(&d, tail + 26 + 8 * i, 8);
            shape.push_back((size_t)d);
          }
          // zero copy: the tensor is the slab block; when its last reference dies the block is queued for release
          // behind an event on the consumer's current stream
          std::shared_ptr<Impl> ip = impl_;          // keeps the bookkeeping alive for tensors that outlive the communicator
          const uint64_t off = b;
          const Device d = land_;
          std::shared_ptr<void> owner(static_cast<void*>(I.dptr + off), [ip, off, d](void*) {
            cudaEvent_t ev = nullptr;
            if (ip->dptr) {                          // slab still registered: release behind the consumer's stream
              d.set_current();
              if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) cudaEventRecord(ev, current_flow(d).stream());
            }
            std::lock_guard<std::mutex> g(ip->mu);
            if (ip->dptr) ip->released.emplace_back(ev, off);
            else if (ev) cudaEventDestroy(ev);
          });
          Message m(static_cast<CommandType>(cmd), Job{make_tensor_view(I.dptr + off, (DType_t)dt, shape, land_, owner), mb});
          m.sender = from;
          {
            std::lock_guard<std::mutex> g(I.mu);
            ++I.st.device_rendezvous_received;
          }
          enqueue(std::move(m));
        }
      } catch (const std::exception& e) {
        std::cerr << "[" << name_ << "] dropped a malformed frame from " << from << ": " << e.what() << std::endl;
      }
    }
  });
}

void RoCECommunicator::stop() {
  if (!running_.exchange(false)) return;
  if (pump_.joinable()) pump_.join();
  bus_->stop();
  Impl& I = *impl_;
  {
    std::lock_guard<std::mutex> g(I.mu);
    for (auto& ps : I.peer_shm) ps.second.close();
    I.peer_shm.clear();
    I.shm.close();
  }
  if (I.dptr) {
    land_.set_current();
    for (auto& o : I.opened) cudaIpcCloseMemHandle(o.second);
    I.opened.clear();
    std::lock_guard<std::mutex> g(I.mu);
    for (auto& r : I.released)
      if (r.first) cudaEventDestroy(r.first);
    I.released.clear();
    if (I.copy_stream) cudaStreamDestroy(I.copy_stream);
    I.copy_stream = nullptr;
    cudaFree(I.dptr);      // tensors that still view the slab must not be USED after the communicator stopped
    I.dptr = nullptr;
  }
}

void RoCECommunicator::send_remote(const Message& m, const Endpoint& to) {
  const std::string key = to.host + ":" + std::to_string(to.port);
  if (!bus_->has_peer(key)) bus_->connect(to.host, to.port);
  Impl& I = *impl_;
  const int64_t t0 = Profiler::now_ns();
  if (m.kind == Message::JOB && m.job.data && m.job.data->is_gpu() && m.job.data->nbytes() > I.inline_bytes) {
    // device rendezvous: announce, the pump thread does the write when the peer grants a block
    current_flow(m.job.data->device()).synchronize();            // the producer of the tensor has finished
    uint64_t id;
    {
      std::lock_guard<std::mutex> g(I.mu);
      id = I.next_id++;
      I.pending[id] = Impl::Pending{m, to};
    }
    const std::string t = I.token + "|" + std::to_string((long long)getpid());
    const auto f = frame(K_GPREPARE, id, m.job.data->nbytes(), t.data(), t.size());
    bus_->send(key, f.data(), f.size());
    GlobalProfiler::add_event(EventType::COMMUNICATION, t0, Profiler::now_ns(), std::string("announce ") + command_name(m.command), name_);
    return;
  }
  std::vector<uint8_t> body = m.serialize();
  if (body.size() > I.inline_bytes && I.shm.base) {
    // large host message: announce it; the pump thread writes it into the block the peer grants (or falls back inline)
    uint64_t id;
    const uint64_t len = body.size();
    {
      std::lock_guard<std::mutex> g(I.mu);
      id = I.next_id++;
      I.pending_host[id] = std::make_pair(std::move(body), key);
    }
    const auto f = frame(K_PREPARE, id, len, I.token.data(), I.token.size());
    bus_->send(key, f.data(), f.size());
    GlobalProfiler::add_event(EventType::COMMUNICATION, t0, Profiler::now_ns(), std::string("announce ") + command_name(m.command), name_);
    return;
  }
  const auto f = frame(K_INLINE, 0, body.size(), body.data(), body.size());
  bus_->send(key, f.data(), f.size());
  {
    std::lock_guard<std::mutex> g(I.mu);
    ++I.st.inline_sent;
  }
  GlobalProfiler::add_event(EventType::COMMUNICATION, t0, Profiler::now_ns(), std::string("send ") + command_name(m.command), name_);
}

}  // namespace tnn
