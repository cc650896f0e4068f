// Profiler: host wall-clock spans {type, start, end, name, source}, mergeable across ranks with start-time re-basing
// and serialisable as a message payload (reference: include/profiling/event.hpp:11-62, profiler.hpp:13-157).
// Byte-compatible with the Python host's ProfilerCore (tnn_b200/csrc/runtime/runtime.cpp).
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <mutex>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost one predictable branch unless a tool is attached

#include "runtime/archiver.hpp"

namespace tnn {

enum class EventType : int { COMPUTE = 0, COMMUNICATION = 1, OTHER = 2 };
struct Event_t {
  EventType type;
  int64_t start_ns, end_ns;
  std::string name, source;
};

class Profiler {
 public:
  Profiler() = default;
  // copies take the SOURCE's lock: the global profiler is copied into a PROFILING_REPORTED message while other workers of
  // the same process keep adding events (found by tools/sanitize_cpp_host.sh)
  Profiler(const Profiler& o) {
    std::lock_guard<std::mutex> g(o.mu_);
    start_ns_ = o.start_ns_;
    events_ = o.events_;
  }
  Profiler& operator=(const Profiler& o) {
    if (this != &o) {
      std::vector<Event_t> ev;
      int64_t st;
      {
        std::lock_guard<std::mutex> g(o.mu_);
        ev = o.events_;
        st = o.start_ns_;
      }
      std::lock_guard<std::mutex> g(mu_);
      start_ns_ = st;
      events_ = std::move(ev);
    }
    return *this;
  }
  static int64_t now_ns() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now().time_since_epoch()).count(); }
  void init_start_time() { start_ns_ = now_ns(); }
  int64_t start_ns() const { return start_ns_; }
  void add_event(EventType t, int64_t s, int64_t e, const std::string& name, const std::string& source) {
    std::lock_guard<std::mutex> g(mu_);
    events_.push_back(Event_t{t, s, e, name, source});
  }
  void merge(const Profiler& o) {                 // re-base the other profiler's events by the start-time difference
    std::lock_guard<std::mutex> g(mu_);
    const int64_t delta = start_ns_ - o.start_ns_;
    for (const auto& ev : o.events_) events_.push_back(Event_t{ev.type, ev.start_ns + delta, ev.end_ns + delta, ev.name, ev.source});
  }
  void sort() {
    std::lock_guard<std::mutex> g(mu_);
    std::sort(events_.begin(), events_.end(), [](const Event_t& a, const Event_t& b) { return a.start_ns < b.start_ns; });
  }
  void clear() {
    std::lock_guard<std::mutex> g(mu_);
    events_.clear();
  }
  std::vector<Event_t> events() const {
    std::lock_guard<std::mutex> g(mu_);
    return events_;
  }
  std::string serialize() const {
    std::lock_guard<std::mutex> g(mu_);
    Writer w;
    w.prim<int64_t>(start_ns_);
    w.prim<uint64_t>(events_.size());
    for (const auto& e : events_) {
      w.prim<uint8_t>((uint8_t)e.type);
      w.prim<int64_t>(e.start_ns);
      w.prim<int64_t>(e.end_ns);
      w.str(e.name);
      w.str(e.source);
    }
    return std::string(reinterpret_cast<const char*>(w.buffer().data()), w.size());
  }
  static Profiler deserialize(const std::string& b) {
    Profiler p;
    Reader r(reinterpret_cast<const uint8_t*>(b.data()), b.size());
    p.start_ns_ = r.prim<int64_t>();
    const uint64_t n = r.prim<uint64_t>();
    for (uint64_t i = 0; i < n; ++i) {
      Event_t e;
      e.type = (EventType)r.prim<uint8_t>();
      e.start_ns = r.prim<int64_t>();
      e.end_ns = r.prim<int64_t>();
      e.name = r.str();
      e.source = r.str();
      p.events_.push_back(std::move(e));
    }
    return p;
  }
  // one line per event, the format visualizers/visualize_profiler.py parses
  void log(const std::string& path) const {
    std::ofstream f(path);
    static const char* names[] = {"COMPUTE", "COMMUNICATION", "OTHER"};
    for (const auto& e : events()) {
      char b[512];
      snprintf(b, sizeof b, "[%s] %s %s start=%.3fms end=%.3fms dur=%.3fms\n", e.source.c_str(), names[(int)e.type], e.name.c_str(),
               (e.start_ns - start_ns_) / 1e6, (e.end_ns - start_ns_) / 1e6, (e.end_ns - e.start_ns) / 1e6);
      f << b;
    }
  }

 private:
  int64_t start_ns_ = 0;
  std::vector<Event_t> events_;
  mutable std::mutex mu_;
};

// NVTX range per layer / training phase (SURVEY 5.1 target: the ranges show up in Nsight Systems / ncu --nvtx).
// Off by default; TNN_NVTX=1 turns them on (read once).
struct NvtxRange {
  static bool enabled() {
    static const bool on = [] { const char* e = getenv("TNN_NVTX"); return e && atoi(e) != 0; }();
    return on;
  }
  explicit NvtxRange(const std::string& name) : active_(enabled()) {
    if (active_) nvtxRangePushA(name.c_str());
  }
  NvtxRange(const std::string& prefix, const std::string& name) : active_(enabled()) {
    if (active_) nvtxRangePushA((prefix + name).c_str());
  }
  ~NvtxRange() {
    if (active_) nvtxRangePop();
  }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;

 private:
  bool active_;
};

// Per-layer timing of the single-process trainer (TrainingConfig::profiler_type / print_layer_profiling; the reference
// declares both, include/nn/train.hpp:31,54-56, and leaves the timing unimplemented).  Sequential records the wall time of
// every child's forward / backward (LayerTimer, layer.hpp: the layer's flow is synchronised at the end of the scope, so a
// profiled run executes eagerly and serialised -- like the reference's Debug build, src/nn/layer.cpp:49-51).  Times of
// blocks include their children.  NORMAL: the table is printed and cleared at every progress interval; CUMULATIVE: it
// accumulates over the run and is printed once at the end.
class LayerProfiler {
 public:
  struct Row {
    std::string name;
    int64_t fwd_ns = 0, bwd_ns = 0;
    uint64_t fwd_calls = 0, bwd_calls = 0;
  };
  static LayerProfiler& instance() {
    static LayerProfiler* p = new LayerProfiler();
    return *p;
  }
  void enable(bool on) { enabled_.store(on, std::memory_order_relaxed); }
  bool enabled() const { return enabled_.load(std::memory_order_relaxed); }
  void add(const std::string& layer, bool backward, int64_t ns) {
    std::lock_guard<std::mutex> g(mu_);
    size_t i = 0;
    for (; i < rows_.size(); ++i)
      if (rows_[i].name == layer) break;
    if (i == rows_.size()) {
      rows_.emplace_back();
      rows_.back().name = layer;
    }
    Row& r = rows_[i];
    if (backward) {
      r.bwd_ns += ns;
      ++r.bwd_calls;
    } else {
      r.fwd_ns += ns;
      ++r.fwd_calls;
    }
  }
  std::vector<Row> rows() const {                       // in first-execution order
    std::lock_guard<std::mutex> g(mu_);
    return rows_;
  }
  void reset() {
    std::lock_guard<std::mutex> g(mu_);
    rows_.clear();
  }
  std::string table(const std::string& title) const {
    const std::vector<Row> rs = rows();
    size_t w = 5;
    for (const auto& r : rs) w = std::max(w, r.name.size());
    std::string out = title + "\n";
    char b[512];
    snprintf(b, sizeof b, "  %-*s %12s %12s %9s %9s %12s %12s\n", (int)w, "layer", "fwd ms", "bwd ms", "fwd calls", "bwd calls", "fwd ms/call", "bwd ms/call");
    out += b;
    double tf = 0, tb = 0;
    for (const auto& r : rs) {
      const double f = r.fwd_ns / 1e6, k = r.bwd_ns / 1e6;
      snprintf(b, sizeof b, "  %-*s %12.3f %12.3f %9llu %9llu %12.4f %12.4f\n", (int)w, r.name.c_str(), f, k, (unsigned long long)r.fwd_calls, (unsigned long long)r.bwd_calls,
               r.fwd_calls ? f / (double)r.fwd_calls : 0.0, r.bwd_calls ? k / (double)r.bwd_calls : 0.0);
      out += b;
      tf += f;
      tb += k;
    }
    snprintf(b, sizeof b, "  %-*s %12.3f %12.3f   (blocks include their children)", (int)w, "sum", tf, tb);
    return out + b;
  }

 private:
  std::atomic<bool> enabled_{false};
  mutable std::mutex mu_;
  std::vector<Row> rows_;
};

class GlobalProfiler {
 public:
  static Profiler& get_profiler() {
    static Profiler* p = [] { auto* q = new Profiler(); q->init_start_time(); return q; }();
    return *p;
  }
  static void init_start_time() { get_profiler().init_start_time(); }
  static void add_event(EventType t, int64_t s, int64_t e, const std::string& name, const std::string& source) { get_profiler().add_event(t, s, e, name, source); }
};

}  // namespace tnn
