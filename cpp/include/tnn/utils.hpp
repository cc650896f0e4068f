// Small host utilities of the C++ API (reference: include/threading/thread_handler.hpp:37-140, include/utils/misc.hpp,
// include/logging/logger.hpp, include/common/concurrent_map.hpp):
//   parallel_for / parallel_for_2d / ThreadWrapper   intra-op CPU threading over the shared pool (csrc/runtime/cpu_math.hpp)
//   ConcurrentMap<K, V>                              mutex-guarded map (emplace / try_get / erase / contains / snapshot)
//   Logger + StepLog                                 levelled console / file logger, JSON-lines metric log
//   get_shape_hash, hash_combine, benchmark()        misc helpers
#pragma once
#include <chrono>
#include <cstdio>
#include <fstream>
#include <functional>
#include <map>
#include <mutex>
#include <optional>
#include <sstream>
#include <cstdlib>
#include <string>
#include <vector>

#include "runtime/cpu_math.hpp"
#include "tnn/json.hpp"

namespace tnn {

// ---- threading
inline int& thread_limit() {
  static int n = 0;        // 0 = pool default (hardware concurrency)
  return n;
}
// fn(i) for i in [begin, end); chunks of `grain` indices per task
inline void parallel_for(int64_t begin, int64_t end, const std::function<void(int64_t)>& fn, int64_t grain = 1) {
  if (end <= begin) return;
  global_pool(thread_limit()).parallel_for(end - begin, grain, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) fn(begin + i);
  });
}
// fn(i, j) over [0, n0) x [0, n1), rows of the outer index per task
inline void parallel_for_2d(int64_t n0, int64_t n1, const std::function<void(int64_t, int64_t)>& fn) {
  parallel_for(0, n0, [&](int64_t i) {
    for (int64_t j = 0; j < n1; ++j) fn(i, j);
  });
}
// run `fn` with the intra-op thread count limited to `num_threads` (reference ThreadWrapper / task_arena)
class ThreadWrapper {
 public:
  explicit ThreadWrapper(int num_threads) : n_(num_threads) {}
  template <class F>
  auto execute(F&& fn) -> decltype(fn()) {
    const int saved = thread_limit();
    thread_limit() = n_;
    struct Restore {
      int v;
      ~Restore() { thread_limit() = v; }
    } r{saved};
    return fn();
  }

 private:
  int n_;
};

// ---- concurrent map
template <class K, class V>
class ConcurrentMap {
 public:
  bool emplace(const K& k, V v) {
    std::lock_guard<std::mutex> g(mu_);
    return map_.emplace(k, std::move(v)).second;
  }
  void set(const K& k, V v) {
    std::lock_guard<std::mutex> g(mu_);
    map_[k] = std::move(v);
  }
  std::optional<V> try_get(const K& k) const {
    std::lock_guard<std::mutex> g(mu_);
    auto it = map_.find(k);
    if (it == map_.end()) return std::nullopt;
    return it->second;
  }
  bool contains(const K& k) const {
    std::lock_guard<std::mutex> g(mu_);
    return map_.count(k) != 0;
  }
  bool erase(const K& k) {
    std::lock_guard<std::mutex> g(mu_);
    return map_.erase(k) != 0;
  }
  size_t size() const {
    std::lock_guard<std::mutex> g(mu_);
    return map_.size();
  }
  std::map<K, V> snapshot() const {
    std::lock_guard<std::mutex> g(mu_);
    return map_;
  }
  // read-modify-write under the lock
  template <class F>
  void update(const K& k, F&& f) {
    std::lock_guard<std::mutex> g(mu_);
    f(map_[k]);
  }

 private:
  mutable std::mutex mu_;
  std::map<K, V> map_;
};

// ---- logging
enum class LogLevel : int { TRACE = 0, DEBUG, INFO, WARN, ERROR, OFF };
class Logger {
 public:
  explicit Logger(std::string name = "tnn", LogLevel level = LogLevel::INFO) : name_(std::move(name)), level_(level) {}
  // the process-wide logger starts at the level named by LOG_LEVEL (TRACE | DEBUG | INFO | WARN | ERROR; the key of the
  // reference's .env.example that none of its C++ reads), INFO otherwise
  static Logger& global() {
    static Logger* l = new Logger("tnn", level_from_string(std::getenv("LOG_LEVEL"), LogLevel::INFO));
    return *l;
  }
  static LogLevel level_from_string(const char* s, LogLevel fallback) {
    if (!s) return fallback;
    std::string u(s);
    for (auto& c : u) c = (char)toupper((unsigned char)c);
    if (u == "TRACE") return LogLevel::TRACE;
    if (u == "DEBUG") return LogLevel::DEBUG;
    if (u == "INFO") return LogLevel::INFO;
    if (u == "WARN" || u == "WARNING") return LogLevel::WARN;
    if (u == "ERROR" || u == "ERR") return LogLevel::ERROR;
    return fallback;
  }
  void set_level(LogLevel l) { level_ = l; }
  LogLevel level() const { return level_; }
  // also append every line to `path`
  void set_file(const std::string& path) {
    std::lock_guard<std::mutex> g(mu_);
    file_.close();
    file_.open(path, std::ios::app);
  }
  void log(LogLevel l, const std::string& msg) {
    if ((int)l < (int)level_) return;
    static const char* tag[] = {"trace", "debug", "info", "warn", "error"};
    const auto now = std::chrono::system_clock::now();
    const std::time_t t = std::chrono::system_clock::to_time_t(now);
    char ts[32];
    std::strftime(ts, sizeof(ts), "%H:%M:%S", std::localtime(&t));
    std::lock_guard<std::mutex> g(mu_);
    std::ostringstream os;
    os << '[' << ts << "] [" << name_ << "] [" << tag[(int)l] << "] " << msg << '\n';
    std::fputs(os.str().c_str(), (int)l >= (int)LogLevel::WARN ? stderr : stdout);
    if (file_.is_open()) {
      file_ << os.str();
      file_.flush();
    }
  }
  void trace(const std::string& m) { log(LogLevel::TRACE, m); }
  void debug(const std::string& m) { log(LogLevel::DEBUG, m); }
  void info(const std::string& m) { log(LogLevel::INFO, m); }
  void warn(const std::string& m) { log(LogLevel::WARN, m); }
  void error(const std::string& m) { log(LogLevel::ERROR, m); }

 private:
  std::string name_;
  LogLevel level_;
  std::mutex mu_;
  std::ofstream file_;
};
// one JSON object per line: {"step": n, "loss": ..., ...}
class StepLog {
 public:
  explicit StepLog(const std::string& path) : f_(path, std::ios::app) {}
  void write(const Json& record) {
    std::lock_guard<std::mutex> g(mu_);
    f_ << record.dump() << '\n';
    f_.flush();
  }
  bool ok() const { return f_.good(); }

 private:
  std::mutex mu_;
  std::ofstream f_;
};

// ---- misc
inline void hash_combine(uint64_t& seed, uint64_t v) { seed ^= v + 0x9E3779B97F4A7C15ull + (seed << 6) + (seed >> 2); }
inline uint64_t get_shape_hash(const std::vector<size_t>& shape) {
  uint64_t h = shape.size();
  for (size_t d : shape) hash_combine(h, (uint64_t)d);
  return h;
}
// wall-clock milliseconds per call after `warmup` calls
inline double benchmark(const std::function<void()>& fn, int iters = 10, int warmup = 2) {
  for (int i = 0; i < warmup; ++i) fn();
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iters; ++i) fn();
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / std::max(1, iters);
}

}  // namespace tnn
