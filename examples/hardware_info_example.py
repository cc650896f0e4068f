#!/usr/bin/env python
"""Host inventory, NUMA map, JSON dump and a short utilisation log (reference:
examples/hardware_info_example.cpp:14-110: writes ./temp/hardware_info.json and ./logs/cpu_status.csv)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tnn_b200.utils.hardware_info import HardwareInfo, ThreadAffinity, get_memory_usage_kb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=5)
    ap.add_argument("--interval", type=float, default=0.2)
    ap.add_argument("--out-dir", default=".")
    a = ap.parse_args()
    print("=== Hardware Information ===")
    hw = HardwareInfo()
    print(hw.summary())
    os.makedirs(os.path.join(a.out_dir, "temp"), exist_ok=True)
    os.makedirs(os.path.join(a.out_dir, "logs"), exist_ok=True)
    hw.update_dynamic_info()
    with open(os.path.join(a.out_dir, "temp", "hardware_info.json"), "w") as f:
        f.write(hw.to_json())
    print("CPU specifications written to temp/hardware_info.json")
    print("=== Distributed Computing Recommendations ===")
    print(f"  worker threads per process: {max(1, hw.physical_cores // max(1, len(hw.gpus) or 1))}; "
          f"affinity now: {len(ThreadAffinity.current())} CPUs")
    print("NUMA topology:")
    for node, cpus in hw.numa_topology().items():
        print(f"  NUMA Node {node}: {', '.join(map(str, cpus[:16]))}{' ...' if len(cpus) > 16 else ''}")
    print("=== JSON Output (displayed) ===")
    print(hw.to_json())
    print("=== Dynamic Monitoring ===")
    path = os.path.join(a.out_dir, "logs", "cpu_status.csv")
    with open(path, "w") as f:
        f.write("time_s,cpu_util_pct,cpu_mhz,temp_c,available_ram_mb,rss_mb\n")
        t0 = time.time()
        for _ in range(a.samples):
            time.sleep(a.interval)
            hw.update_dynamic_info()
            f.write(f"{time.time() - t0:.2f},{hw.cpu_utilization:.1f},{hw.cpu_mhz:.0f},{hw.temperature_c:.1f},"
                    f"{hw.available_ram_kb // 1024},{get_memory_usage_kb() // 1024}\n")
            print(f"Update: CPU: {hw.cpu_utilization:.1f}%, Temp: {hw.temperature_c:.1f}C, Freq: {hw.cpu_mhz:.0f} MHz")
    print(f"CPU status data written to {path}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
