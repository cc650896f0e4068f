// hardware_info_example: host inventory, NUMA map, JSON dump and a short utilisation log (reference:
// examples/hardware_info_example.cpp:14-110: writes ./temp/hardware_info.json and ./logs/cpu_status.csv).
//   hardware_info_example [--samples N] [--interval-ms M] [--out-dir DIR]
#include <sys/stat.h>

#include <chrono>
#include <cstring>
#include <iostream>
#include <thread>

#include "tnn/hardware_info.hpp"

using namespace tnn;

int main(int argc, char** argv) {
  int samples = 5, interval_ms = 200;
  std::string out_dir = ".";
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--samples") && i + 1 < argc) samples = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--interval-ms") && i + 1 < argc) interval_ms = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--out-dir") && i + 1 < argc) out_dir = argv[++i];
  }
  std::cout << "=== Hardware Information ===" << std::endl;
  HardwareInfo hw;
  std::cout << hw.summary();
  mkdir((out_dir + "/temp").c_str(), 0755);
  mkdir((out_dir + "/logs").c_str(), 0755);
  hw.update_dynamic_info();
  {
    std::ofstream f(out_dir + "/temp/hardware_info.json");
    f << hw.to_json().dump(2) << std::endl;
  }
  std::cout << "CPU specifications written to temp/hardware_info.json" << std::endl;
  std::cout << "=== Distributed Computing Recommendations ===" << std::endl;
  std::cout << "  worker threads per process: " << std::max<size_t>(1, (size_t)hw.physical_cores / std::max<size_t>(1, hw.gpus.size()))
            << "; affinity now: " << ThreadAffinity::current().size() << " CPUs" << std::endl;
  std::cout << "NUMA topology:" << std::endl;
  for (auto& kv : hw.numa_topology()) {
    std::cout << "  NUMA Node " << kv.first << ":";
    for (size_t i = 0; i < kv.second.size() && i < 16; ++i) std::cout << ' ' << kv.second[i];
    std::cout << (kv.second.size() > 16 ? " ..." : "") << std::endl;
  }
  std::cout << "=== JSON Output (displayed) ===\n" << hw.to_json().dump(2) << std::endl;
  std::cout << "=== Dynamic Monitoring ===" << std::endl;
  std::ofstream csv(out_dir + "/logs/cpu_status.csv");
  csv << "time_s,cpu_util_pct,cpu_mhz,temp_c,available_ram_mb,rss_mb\n";
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < samples; ++i) {
    std::this_thread::sleep_for(std::chrono::milliseconds(interval_ms));
    hw.update_dynamic_info();
    const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    csv << t << ',' << hw.cpu_util_pct << ',' << hw.cur_mhz << ',' << hw.temp_c << ',' << hw.available_ram_kb / 1024 << ',' << get_memory_usage_kb() / 1024 << '\n';
    std::cout << "  t=" << t << "s util " << hw.cpu_util_pct << "% " << hw.cur_mhz << " MHz, available RAM " << hw.available_ram_kb / 1024 << " MiB" << std::endl;
  }
  std::cout << "log written to logs/cpu_status.csv" << std::endl;
  return 0;
}
