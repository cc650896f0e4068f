// TCP message bus: the control-plane / multi-node transport of the coordinator-worker runtime.
//
// Capability parity with the reference TCP communicator (include/distributed/
// tcp_communicator.hpp:59-474, tcp_channel.hpp, chunking.hpp): several sockets per peer, messages
// cut into <= 4 MiB packets that are round-robined over those sockets, reassembly by message
// serial id, an identity handshake carrying the sender's listening port, TCP_NODELAY, and a
// receive side that hands out complete messages in CommandType priority order (the first two
// bytes of every serialised Message are its CommandType; lower value = higher priority,
// message_map.hpp:16-89).  Implementation is plain POSIX sockets + std::thread (the reference
// uses Asio coroutines, which are not available offline); one reader thread per socket.
//
// On a single 8xB200 node tensors do NOT travel through this bus: stage activations move with
// NCCL p2p / peer copies (parallel/pipeline.py) and this bus carries control messages only.  It
// still moves full tensor payloads (host staged, like the reference) when endpoints are on
// different hosts or in the CPU tests.
#pragma once
#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "archiver.hpp"

namespace tnn {

class TcpBus {
 public:
  static constexpr uint32_t kMagic = 0x544E4E42;  // "TNNB"
  static constexpr int kPriorities = 64;

  struct Stats {
    uint64_t bytes_sent = 0, bytes_received = 0, messages_sent = 0, messages_received = 0, packets_sent = 0,
             bad_packets = 0;
  };

  // `bind_host`: interface the listener binds to ("127.0.0.1" by default: the bus is unauthenticated, so it is
  // NOT exposed on every interface unless the caller asks for "0.0.0.0").  `max_message`: upper bound accepted
  // for a peer-declared message length (a peer cannot make this process allocate more than that).
  TcpBus(int listen_port, int sockets_per_peer = 4, size_t max_packet = 4u << 20,
         const std::string& bind_host = "127.0.0.1", uint64_t max_message = 8ull << 30)
      : port_(listen_port), spp_(sockets_per_peer < 1 ? 1 : sockets_per_peer), max_packet_(max_packet),
        bind_host_(bind_host), max_message_(max_message) {}
  ~TcpBus() { stop(); }

  int port() const { return port_; }

  void start() {
    if (running_.exchange(true)) return;
    const int lfd = ::socket(AF_INET, SOCK_STREAM, 0);
    if (lfd < 0) throw std::runtime_error("TcpBus: socket() failed");
    lfd_ = lfd;
    int one = 1;
    setsockopt(lfd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in addr{};
    addr.sin_family = AF_INET;
    if (bind_host_.empty() || bind_host_ == "0.0.0.0" || bind_host_ == "*") {
      addr.sin_addr.s_addr = htonl(INADDR_ANY);
    } else if (inet_pton(AF_INET, bind_host_ == "localhost" ? "127.0.0.1" : bind_host_.c_str(), &addr.sin_addr) != 1) {
      addrinfo hints{}, *res = nullptr;
      hints.ai_family = AF_INET;
      hints.ai_socktype = SOCK_STREAM;
      if (getaddrinfo(bind_host_.c_str(), nullptr, &hints, &res) != 0 || res == nullptr) {
        ::close(lfd);
        lfd_ = -1;
        running_ = false;
        throw std::runtime_error("TcpBus: cannot resolve bind host " + bind_host_);
      }
      addr.sin_addr = reinterpret_cast<sockaddr_in*>(res->ai_addr)->sin_addr;
      freeaddrinfo(res);
    }
    addr.sin_port = htons((uint16_t)port_);
    if (::bind(lfd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0) {
      ::close(lfd);
      lfd_ = -1;
      running_ = false;
      throw std::runtime_error("TcpBus: bind() failed on port " + std::to_string(port_));
    }
    if (port_ == 0) {
      socklen_t len = sizeof(addr);
      getsockname(lfd, reinterpret_cast<sockaddr*>(&addr), &len);
      port_ = ntohs(addr.sin_port);
    }
    ::listen(lfd, 64);
    acceptor_ = std::thread([this] { accept_loop(); });
  }

  void stop() {
    if (!running_.exchange(false)) return;
    const int lfd = lfd_.load();
    if (lfd >= 0) ::shutdown(lfd, SHUT_RDWR);
    if (acceptor_.joinable()) acceptor_.join();   // the acceptor polls running_; close only after it left
    if (lfd >= 0) {
      ::close(lfd);
      lfd_ = -1;
    }
    std::vector<std::thread> readers;
    {
      std::lock_guard<std::mutex> g(mu_);
      for (auto& kv : peers_)
        for (auto& ch : kv.second->chans) {
          if (ch->fd >= 0) ::shutdown(ch->fd, SHUT_RDWR);
        }
      readers.swap(readers_);
    }
    for (auto& t : readers)
      if (t.joinable()) t.join();
    {
      std::lock_guard<std::mutex> g(mu_);
      for (auto& kv : peers_)
        for (auto& ch : kv.second->chans)
          if (ch->fd >= 0) {
            ::close(ch->fd);
            ch->fd = -1;
          }
      peers_.clear();
    }
    cv_.notify_all();
  }

  // Opens `sockets_per_peer` connections to host:port and identifies this bus by its own
  // listening port, so that the peer can address replies to "our_host:our_port".
  void connect(const std::string& host, int port, int timeout_ms = 10000) {
    const std::string key = host + ":" + std::to_string(port);
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = peers_.find(key);
      if (it != peers_.end() && !it->second->chans.empty()) return;
    }
    for (int i = 0; i < spp_; ++i) {
      int fd = dial(host, port, timeout_ms);
      uint32_t hello[3] = {kMagic, (uint32_t)port_, (uint32_t)i};
      write_all(fd, hello, sizeof(hello));
      add_channel(key, fd);
    }
  }

  bool has_peer(const std::string& key) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = peers_.find(key);
    return it != peers_.end() && !it->second->chans.empty();
  }

  std::vector<std::string> peers() {
    std::lock_guard<std::mutex> g(mu_);
    std::vector<std::string> out;
    for (auto& kv : peers_)
      if (!kv.second->chans.empty()) out.push_back(kv.first);
    return out;
  }

  // Slices one serialised message into packets and writes them round-robin over the peer's sockets.
  void send(const std::string& key, const uint8_t* data, size_t n) {
    std::shared_ptr<Peer> peer;
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = peers_.find(key);
      if (it == peers_.end() || it->second->chans.empty()) throw std::runtime_error("TcpBus: no connection to " + key);
      peer = it->second;
    }
    const uint64_t serial = next_serial_.fetch_add(1);
    const uint32_t total = (uint32_t)std::max<size_t>(1, (n + max_packet_ - 1) / max_packet_);
    for (uint32_t i = 0; i < total; ++i) {
      const size_t off = (size_t)i * max_packet_;
      const size_t len = std::min(max_packet_, n - std::min(n, off));
      PacketHeader h;
      h.type = 0;
      h.endianness = host_is_little_endian() ? 1 : 0;
      h.packet_length = len;
      h.msg_length = n;
      h.msg_serial_id = serial;
      h.packet_offset = i;
      h.total_packets = total;
      Writer w;
      h.write(w);
      auto& ch = peer->chans[(peer->rr.fetch_add(1)) % peer->chans.size()];
      std::lock_guard<std::mutex> wg(ch->wmu);
      write_all(ch->fd, w.buffer().data(), w.size());
      if (len) write_all(ch->fd, data + off, len);
      stats_packets_.fetch_add(1);
    }
    stats_bytes_sent_.fetch_add(n);
    stats_msgs_sent_.fetch_add(1);
  }

  // Highest-priority complete message, or false after timeout_ms (negative = wait forever).
  bool recv(int timeout_ms, std::string* from, std::vector<uint8_t>* msg) {
    std::unique_lock<std::mutex> lk(qmu_);
    auto ready = [this] { return queued_ > 0 || !running_; };
    if (timeout_ms < 0)
      cv_.wait(lk, ready);
    else if (!cv_.wait_for(lk, std::chrono::milliseconds(timeout_ms), ready))
      return false;
    if (queued_ == 0) return false;
    for (int p = 0; p < kPriorities; ++p) {
      auto& q = queues_[p];
      if (!q.empty()) {
        *from = std::move(q.front().first);
        *msg = std::move(q.front().second);
        q.pop_front();
        --queued_;
        return true;
      }
    }
    return false;
  }

  Stats stats() const {
    Stats s;
    s.bytes_sent = stats_bytes_sent_;
    s.bytes_received = stats_bytes_recv_;
    s.messages_sent = stats_msgs_sent_;
    s.messages_received = stats_msgs_recv_;
    s.packets_sent = stats_packets_;
    s.bad_packets = stats_bad_packets_;
    return s;
  }

 private:
  struct Channel {
    int fd = -1;
    std::mutex wmu;
  };
  struct Pending {
    std::vector<uint8_t> buf;
    uint64_t len = 0;                 // declared message length (fixed for the life of the serial)
    std::vector<uint8_t> seen;        // one flag per slice: duplicates are rejected
    uint32_t got = 0;
  };
  struct Peer {
    std::vector<std::shared_ptr<Channel>> chans;
    std::atomic<uint32_t> rr{0};
    std::mutex amu;
    std::map<uint64_t, Pending> assembling;
  };

  static void write_all(int fd, const void* p, size_t n) {
    const uint8_t* b = reinterpret_cast<const uint8_t*>(p);
    while (n > 0) {
      ssize_t k = ::send(fd, b, n, MSG_NOSIGNAL);
      if (k <= 0) {
        if (k < 0 && (errno == EINTR || errno == EAGAIN)) continue;
        throw std::runtime_error("TcpBus: send failed");
      }
      b += k;
      n -= (size_t)k;
    }
  }
  static bool read_all(int fd, void* p, size_t n) {
    uint8_t* b = reinterpret_cast<uint8_t*>(p);
    while (n > 0) {
      ssize_t k = ::recv(fd, b, n, 0);
      if (k == 0) return false;
      if (k < 0) {
        if (errno == EINTR) continue;
        return false;
      }
      b += k;
      n -= (size_t)k;
    }
    return true;
  }
  static void tune(int fd) {
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    int buf = 8 << 20;
    setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &buf, sizeof(buf));
    setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &buf, sizeof(buf));
  }
  static int dial(const std::string& host, int port, int timeout_ms) {
    addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    if (getaddrinfo(host.c_str(), std::to_string(port).c_str(), &hints, &res) != 0 || res == nullptr)
      throw std::runtime_error("TcpBus: cannot resolve " + host);
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms);
    int fd = -1;
    while (true) {
      fd = ::socket(AF_INET, SOCK_STREAM, 0);
      if (fd >= 0 && ::connect(fd, res->ai_addr, res->ai_addrlen) == 0) break;
      if (fd >= 0) ::close(fd);
      fd = -1;
      if (std::chrono::steady_clock::now() > deadline) break;
      std::this_thread::sleep_for(std::chrono::milliseconds(50));  // peer may not be listening yet
    }
    freeaddrinfo(res);
    if (fd < 0) throw std::runtime_error("TcpBus: connect to " + host + ":" + std::to_string(port) + " timed out");
    tune(fd);
    return fd;
  }

  void add_channel(const std::string& key, int fd) {
    auto ch = std::make_shared<Channel>();
    ch->fd = fd;
    std::shared_ptr<Peer> peer;
    {
      std::lock_guard<std::mutex> g(mu_);
      auto& p = peers_[key];
      if (!p) p = std::make_shared<Peer>();
      p->chans.push_back(ch);
      peer = p;
      readers_.emplace_back([this, key, peer, ch] { read_loop(key, peer, ch); });
    }
  }

  void accept_loop() {
    const int lfd = lfd_.load();
    while (running_) try {
      pollfd pfd{lfd, POLLIN, 0};
      int r = ::poll(&pfd, 1, 200);
      if (r <= 0) continue;
      sockaddr_in peer_addr{};
      socklen_t len = sizeof(peer_addr);
      int fd = ::accept(lfd, reinterpret_cast<sockaddr*>(&peer_addr), &len);
      if (fd < 0) continue;
      tune(fd);
      uint32_t hello[3];
      if (!read_all(fd, hello, sizeof(hello)) || hello[0] != kMagic) {
        ::close(fd);
        continue;
      }
      char ip[64];
      inet_ntop(AF_INET, &peer_addr.sin_addr, ip, sizeof(ip));
      add_channel(std::string(ip) + ":" + std::to_string(hello[1]), fd);
    } catch (...) {
      stats_bad_packets_.fetch_add(1);    // e.g. thread creation failed: keep accepting
    }
  }

  // One reader thread per socket.  Everything a peer controls is validated before it sizes or indexes a buffer:
  // message length (<= max_message_), packet count (== ceil(msg / max_packet)), packet offset (< count), packet
  // length (exactly the slice it claims to be) and a serial's length may not change once its buffer exists (a second
  // packet with another msg_length would reallocate under a concurrent reader).  Any violation, I/O error or
  // allocation failure drops this channel; nothing thrown here may escape the thread (std::terminate).
  void read_loop(const std::string& key, std::shared_ptr<Peer> peer, std::shared_ptr<Channel> ch) {
    try {
      read_loop_impl(key, peer, ch);
    } catch (...) {
      stats_bad_packets_.fetch_add(1);
    }
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = peers_.find(key);
      if (it != peers_.end()) {
        auto& v = it->second->chans;
        for (size_t i = 0; i < v.size(); ++i)
          if (v[i] == ch) {
            v.erase(v.begin() + i);
            break;
          }
      }
    }
  }

  void read_loop_impl(const std::string& key, std::shared_ptr<Peer>& peer, std::shared_ptr<Channel>& ch) {
    std::vector<uint8_t> hb(PacketHeader::kBytes);
    while (running_) {
      if (!read_all(ch->fd, hb.data(), hb.size())) break;
      PacketHeader h = PacketHeader::read(hb.data(), hb.size());
      if (h.version != 1 || h.total_packets == 0 || h.packet_length > max_packet_ || h.msg_length > max_message_) {
        stats_bad_packets_.fetch_add(1);
        break;
      }
      const uint64_t want_packets = std::max<uint64_t>(1, (h.msg_length + max_packet_ - 1) / max_packet_);
      if (h.total_packets != want_packets || h.packet_offset >= h.total_packets) {
        stats_bad_packets_.fetch_add(1);
        break;
      }
      const uint64_t off = (uint64_t)h.packet_offset * max_packet_;
      const uint64_t want_len = std::min<uint64_t>(max_packet_, h.msg_length - std::min<uint64_t>(h.msg_length, off));
      if (h.packet_length != want_len) {
        stats_bad_packets_.fetch_add(1);
        break;
      }
      std::vector<uint8_t>* dst;
      {
        std::lock_guard<std::mutex> g(peer->amu);
        auto it = peer->assembling.find(h.msg_serial_id);
        if (it == peer->assembling.end()) {
          if (peer->assembling.size() >= kMaxAssembling) {
            stats_bad_packets_.fetch_add(1);
            break;
          }
          Pending pd;
          pd.buf.resize(h.msg_length);
          pd.len = h.msg_length;
          pd.seen.assign(h.total_packets, 0);
          it = peer->assembling.emplace(h.msg_serial_id, std::move(pd)).first;
        } else if (it->second.len != h.msg_length || it->second.seen.size() != h.total_packets) {
          stats_bad_packets_.fetch_add(1);
          break;                                  // size change for a live serial: never reallocate
        }
        if (it->second.seen[h.packet_offset]) {
          stats_bad_packets_.fetch_add(1);
          break;                                  // duplicate slice
        }
        it->second.seen[h.packet_offset] = 1;
        dst = &it->second.buf;
      }
      // distinct packets of one message land in disjoint ranges, so concurrent reader threads
      // (one per socket) can fill the same buffer without holding the lock
      if (h.packet_length && !read_all(ch->fd, dst->data() + off, h.packet_length)) break;
      bool complete = false;
      std::vector<uint8_t> done;
      {
        std::lock_guard<std::mutex> g(peer->amu);
        auto it = peer->assembling.find(h.msg_serial_id);
        if (it == peer->assembling.end()) break;
        if (++it->second.got == h.total_packets) {
          done = std::move(it->second.buf);
          peer->assembling.erase(it);
          complete = true;
        }
      }
      if (complete) {
        int prio = kPriorities - 1;
        if (done.size() >= 2) {
          uint16_t cmd;
          std::memcpy(&cmd, done.data(), 2);
          if (h.endianness != (host_is_little_endian() ? 1 : 0)) cmd = (uint16_t)((cmd >> 8) | (cmd << 8));
          prio = std::min<int>(cmd, kPriorities - 1);
        }
        stats_bytes_recv_.fetch_add(done.size());
        stats_msgs_recv_.fetch_add(1);
        {
          std::lock_guard<std::mutex> g(qmu_);
          queues_[prio].emplace_back(key, std::move(done));
          ++queued_;
        }
        cv_.notify_one();
      }
    }
  }

  static constexpr size_t kMaxAssembling = 4096;   // messages in flight per peer
  int port_;
  int spp_;
  size_t max_packet_;
  std::string bind_host_;
  uint64_t max_message_;
  std::atomic<uint64_t> stats_bad_packets_{0};
  std::atomic<int> lfd_{-1};
  std::atomic<bool> running_{false};
  std::thread acceptor_;
  std::mutex mu_;
  std::map<std::string, std::shared_ptr<Peer>> peers_;
  std::vector<std::thread> readers_;
  std::atomic<uint64_t> next_serial_{1};
  std::mutex qmu_;
  std::condition_variable cv_;
  std::deque<std::pair<std::string, std::vector<uint8_t>>> queues_[kPriorities];
  size_t queued_ = 0;
  std::atomic<uint64_t> stats_bytes_sent_{0}, stats_bytes_recv_{0}, stats_msgs_sent_{0}, stats_msgs_recv_{0},
      stats_packets_{0};
};

}  // namespace tnn
