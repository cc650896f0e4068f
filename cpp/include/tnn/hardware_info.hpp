// Host / GPU inventory (reference: include/utils/hardware_info.hpp:126, src/utils/hardware_info.cpp, thread_affinity.hpp:46,
// misc.hpp get_memory_usage_kb): CPU model / sockets / physical + logical cores / caches / frequency, RAM, NUMA map,
// thermal zones, load, the visible GPUs, JSON dump; ThreadAffinity pins the calling thread.  Linux procfs / sysfs.
#pragma once
#include <cuda_runtime.h>
#include <sched.h>
#include <unistd.h>

#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "tnn/json.hpp"

namespace tnn {

struct GpuInfo {
  int index = 0;
  std::string name;
  int cc_major = 0, cc_minor = 0, sm_count = 0;
  size_t total_mem = 0, l2_bytes = 0, smem_per_sm = 0;
  int clock_khz = 0, mem_bus_bits = 0;
};

inline long get_memory_usage_kb() {               // resident set of this process
  std::ifstream f("/proc/self/status");
  std::string line;
  while (std::getline(f, line))
    if (line.rfind("VmRSS:", 0) == 0) return atol(line.c_str() + 6);
  return 0;
}

class HardwareInfo {
 public:
  HardwareInfo() { refresh_static(); update_dynamic_info(); }
  std::string cpu_model, vendor, arch;
  int sockets = 0, physical_cores = 0, logical_cores = 0;
  double base_mhz = 0, cur_mhz = 0;
  std::map<std::string, std::string> caches;      // "L1d" -> "48K" ...
  size_t total_ram_kb = 0, available_ram_kb = 0;
  double load1 = 0, cpu_util_pct = 0, temp_c = -1;
  std::vector<GpuInfo> gpus;
  bool has_avx2 = false, has_avx512 = false;

  void refresh_static() {
    std::ifstream f("/proc/cpuinfo");
    std::string line;
    std::set<std::pair<int, int>> cores;
    std::set<int> socks;
    int phys = 0, core = 0;
    logical_cores = 0;
    while (std::getline(f, line)) {
      const size_t c = line.find(':');
      if (c == std::string::npos) continue;
      std::string k = line.substr(0, c), v = c + 2 <= line.size() ? line.substr(c + 2) : "";
      while (!k.empty() && (k.back() == '\t' || k.back() == ' ')) k.pop_back();
      if (k == "processor") ++logical_cores;
      else if (k == "model name" && cpu_model.empty()) cpu_model = v;
      else if (k == "vendor_id" && vendor.empty()) vendor = v;
      else if (k == "physical id") { phys = atoi(v.c_str()); socks.insert(phys); }
      else if (k == "core id") { core = atoi(v.c_str()); cores.insert({phys, core}); }
      else if (k == "cpu MHz" && base_mhz == 0) base_mhz = atof(v.c_str());
      else if (k == "flags" && !has_avx2) {
        has_avx2 = v.find(" avx2") != std::string::npos;
        has_avx512 = v.find(" avx512f") != std::string::npos;
      }
    }
    sockets = socks.empty() ? 1 : (int)socks.size();
    physical_cores = cores.empty() ? logical_cores : (int)cores.size();
#if defined(__x86_64__)
    arch = "x86_64";
#elif defined(__aarch64__)
    arch = "aarch64";
#else
    arch = "unknown";
#endif
    for (int i = 0; i < 8; ++i) {
      const std::string base = "/sys/devices/system/cpu/cpu0/cache/index" + std::to_string(i) + "/";
      std::string level = read_line(base + "level"), type = read_line(base + "type"), size = read_line(base + "size");
      if (level.empty()) break;
      caches["L" + level + (type == "Data" ? "d" : type == "Instruction" ? "i" : "")] = size;
    }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { n = 0; cudaGetLastError(); }
    gpus.clear();
    for (int i = 0; i < n; ++i) {
      cudaDeviceProp p{};
      if (cudaGetDeviceProperties(&p, i) != cudaSuccess) continue;
      GpuInfo g;
      g.index = i;
      g.name = p.name;
      g.cc_major = p.major;
      g.cc_minor = p.minor;
      g.sm_count = p.multiProcessorCount;
      g.total_mem = p.totalGlobalMem;
      g.l2_bytes = (size_t)p.l2CacheSize;
      g.smem_per_sm = p.sharedMemPerMultiprocessor;
      g.mem_bus_bits = p.memoryBusWidth;
      cudaDeviceGetAttribute(&g.clock_khz, cudaDevAttrClockRate, i);
      gpus.push_back(g);
    }
  }
  // RAM, load, utilisation since the previous call, current frequency, hottest thermal zone
  void update_dynamic_info() {
    std::ifstream m("/proc/meminfo");
    std::string k;
    size_t v;
    std::string unit;
    while (m >> k >> v) {
      std::getline(m, unit);
      if (k == "MemTotal:") total_ram_kb = v;
      else if (k == "MemAvailable:") available_ram_kb = v;
    }
    std::ifstream l("/proc/loadavg");
    l >> load1;
    std::ifstream s("/proc/stat");
    std::string cpu;
    unsigned long long a[8] = {0};
    s >> cpu;
    for (auto& x : a) s >> x;
    const unsigned long long idle = a[3] + a[4], total = a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7];
    if (prev_total_ && total > prev_total_) cpu_util_pct = 100.0 * (1.0 - (double)(idle - prev_idle_) / (double)(total - prev_total_));
    prev_idle_ = idle;
    prev_total_ = total;
    const std::string f = read_line("/sys/devices/system/cpu/cpu0/cpufreq/scaling_cur_freq");
    cur_mhz = f.empty() ? base_mhz : atof(f.c_str()) / 1000.0;
    temp_c = -1;
    for (int i = 0; i < 16; ++i) {
      const std::string t = read_line("/sys/class/thermal/thermal_zone" + std::to_string(i) + "/temp");
      if (t.empty()) break;
      temp_c = std::max(temp_c, atof(t.c_str()) / 1000.0);
    }
  }
  std::map<int, std::vector<int>> numa_topology() const {
    std::map<int, std::vector<int>> out;
    for (int n = 0; n < 64; ++n) {
      const std::string list = read_line("/sys/devices/system/node/node" + std::to_string(n) + "/cpulist");
      if (list.empty()) break;
      out[n] = parse_cpulist(list);
    }
    if (out.empty())
      for (int i = 0; i < logical_cores; ++i) out[0].push_back(i);
    return out;
  }
  Json to_json() const {
    Json j = Json::object();
    Json cpu = Json::object();
    cpu["model"] = cpu_model;
    cpu["vendor"] = vendor;
    cpu["architecture"] = arch;
    cpu["sockets"] = sockets;
    cpu["physical_cores"] = physical_cores;
    cpu["logical_cores"] = logical_cores;
    cpu["base_mhz"] = base_mhz;
    cpu["current_mhz"] = cur_mhz;
    cpu["avx2"] = has_avx2;
    cpu["avx512"] = has_avx512;
    Json c = Json::object();
    for (auto& kv : caches) c[kv.first] = kv.second;
    cpu["caches"] = c;
    cpu["utilization_pct"] = cpu_util_pct;
    cpu["load1"] = load1;
    cpu["temperature_c"] = temp_c;
    j["cpu"] = cpu;
    Json mem = Json::object();
    mem["total_kb"] = (long long)total_ram_kb;
    mem["available_kb"] = (long long)available_ram_kb;
    mem["process_rss_kb"] = (long long)get_memory_usage_kb();
    j["memory"] = mem;
    Json numa = Json::object();
    for (auto& kv : numa_topology()) numa[std::to_string(kv.first)] = Json::from_vector(kv.second);
    j["numa"] = numa;
    Json gs = Json::array();
    for (const GpuInfo& g : gpus) {
      Json o = Json::object();
      o["index"] = g.index;
      o["name"] = g.name;
      o["compute_capability"] = std::to_string(g.cc_major) + "." + std::to_string(g.cc_minor);
      o["sm_count"] = g.sm_count;
      o["total_memory_bytes"] = (long long)g.total_mem;
      o["l2_bytes"] = (long long)g.l2_bytes;
      o["smem_per_sm"] = (long long)g.smem_per_sm;
      o["clock_khz"] = g.clock_khz;
      o["memory_bus_bits"] = g.mem_bus_bits;
      gs.push_back(o);
    }
    j["gpus"] = gs;
    return j;
  }
  std::string summary() const {
    std::ostringstream os;
    os << "CPU: " << cpu_model << " (" << vendor << ", " << arch << ")\n"
       << "  sockets " << sockets << ", physical cores " << physical_cores << ", logical cores " << logical_cores << ", " << base_mhz << " MHz, AVX2 "
       << (has_avx2 ? "yes" : "no") << ", AVX-512 " << (has_avx512 ? "yes" : "no") << "\n  caches:";
    for (auto& kv : caches) os << ' ' << kv.first << '=' << kv.second;
    os << "\nRAM: " << total_ram_kb / 1024 << " MiB total, " << available_ram_kb / 1024 << " MiB available\n";
    for (const GpuInfo& g : gpus)
      os << "GPU " << g.index << ": " << g.name << " sm_" << g.cc_major << g.cc_minor << ", " << g.sm_count << " SMs, " << (g.total_mem >> 20) << " MiB, L2 "
         << (g.l2_bytes >> 20) << " MiB\n";
    if (gpus.empty()) os << "GPU: none visible\n";
    return os.str();
  }

 private:
  static std::string read_line(const std::string& path) {
    std::ifstream f(path);
    std::string s;
    if (f) std::getline(f, s);
    return s;
  }
  static std::vector<int> parse_cpulist(const std::string& s) {
    std::vector<int> out;
    std::stringstream ss(s);
    std::string part;
    while (std::getline(ss, part, ',')) {
      const size_t d = part.find('-');
      const int a = atoi(part.c_str()), b = d == std::string::npos ? a : atoi(part.c_str() + d + 1);
      for (int i = a; i <= b; ++i) out.push_back(i);
    }
    return out;
  }
  unsigned long long prev_idle_ = 0, prev_total_ = 0;
};

// Pin the calling thread (thread_affinity.hpp:46-120: the reference pins workers to P-cores; here an explicit CPU set)
struct ThreadAffinity {
  static std::vector<int> current() {
    cpu_set_t set;
    CPU_ZERO(&set);
    std::vector<int> out;
    if (sched_getaffinity(0, sizeof(set), &set) == 0)
      for (int i = 0; i < CPU_SETSIZE; ++i)
        if (CPU_ISSET(i, &set)) out.push_back(i);
    return out;
  }
  static bool pin(const std::vector<int>& cpus) {
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c : cpus) CPU_SET(c, &set);
    return sched_setaffinity(0, sizeof(set), &set) == 0;
  }
  static bool pin_to_numa_node(const HardwareInfo& hw, int node) {
    auto topo = hw.numa_topology();
    auto it = topo.find(node);
    return it != topo.end() && pin(it->second);
  }
};

}  // namespace tnn
