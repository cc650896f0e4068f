"""Host / device inventory (reference: include/utils/hardware_info.hpp:126, thread_affinity.hpp:46,
misc.hpp:10-50): CPU model + topology, caches, RAM, GPUs; process memory; thread pinning."""
from __future__ import annotations

import hashlib
import os
import platform


class HardwareInfo:
    def __init__(self):
        self.cpu_model = platform.processor() or ""
        self.logical_cores = os.cpu_count() or 1
        self.physical_cores = self.logical_cores
        self.sockets = 1
        self.caches = {}
        self.total_ram_kb = self.available_ram_kb = 0
        self.gpus = []
        self._probe()

    def _probe(self):
        try:
            with open("/proc/cpuinfo") as f:
                txt = f.read()
            models = [l.split(":", 1)[1].strip() for l in txt.splitlines() if l.startswith("model name")]
            if models:
                self.cpu_model = models[0]
            phys = {l.split(":", 1)[1].strip() for l in txt.splitlines() if l.startswith("physical id")}
            self.sockets = max(1, len(phys))
            cores = {(a, b) for a, b in zip(
                [l.split(":", 1)[1].strip() for l in txt.splitlines() if l.startswith("physical id")],
                [l.split(":", 1)[1].strip() for l in txt.splitlines() if l.startswith("core id")])}
            if cores:
                self.physical_cores = len(cores)
        except OSError:
            pass
        try:
            with open("/proc/meminfo") as f:
                for l in f:
                    if l.startswith("MemTotal"):
                        self.total_ram_kb = int(l.split()[1])
                    elif l.startswith("MemAvailable"):
                        self.available_ram_kb = int(l.split()[1])
        except OSError:
            pass
        base = "/sys/devices/system/cpu/cpu0/cache"
        if os.path.isdir(base):
            for idx in sorted(os.listdir(base)):
                try:
                    lvl = open(f"{base}/{idx}/level").read().strip()
                    typ = open(f"{base}/{idx}/type").read().strip()
                    size = open(f"{base}/{idx}/size").read().strip()
                    self.caches[f"L{lvl}{typ[0].lower() if typ != 'Unified' else ''}"] = size
                except OSError:
                    continue
        try:
            import torch
            if torch.cuda.is_available():
                for i in range(torch.cuda.device_count()):
                    p = torch.cuda.get_device_properties(i)
                    self.gpus.append(dict(index=i, name=p.name, sms=p.multi_processor_count,
                                          memory_gb=round(p.total_memory / 2 ** 30, 1), cc=f"{p.major}.{p.minor}"))
        except Exception:  # noqa: BLE001
            pass

    def summary(self) -> str:
        lines = [f"CPU: {self.cpu_model} ({self.sockets} socket(s), {self.physical_cores} cores / {self.logical_cores} threads)",
                 f"Caches: {self.caches}", f"RAM: {self.total_ram_kb / 2 ** 20:.1f} GiB total, {self.available_ram_kb / 2 ** 20:.1f} GiB available"]
        for g in self.gpus:
            lines.append(f"GPU {g['index']}: {g['name']} sm_{g['cc'].replace('.', '')} {g['sms']} SMs {g['memory_gb']} GiB")
        return "\n".join(lines)


    # ---- topology / dynamic state (hardware_info.hpp: NUMA map, utilisation, temperature, JSON dump)
    @staticmethod
    def _parse_cpulist(txt: str) -> list:
        out = []
        for part in txt.strip().split(","):
            if not part:
                continue
            if "-" in part:
                a, b = part.split("-")
                out.extend(range(int(a), int(b) + 1))
            else:
                out.append(int(part))
        return out

    def numa_topology(self) -> dict:
        """{node id: [cpu ids]} from sysfs; one node holding every CPU when the kernel exposes none."""
        base = "/sys/devices/system/node"
        nodes = {}
        if os.path.isdir(base):
            for d in sorted(os.listdir(base)):
                if d.startswith("node") and d[4:].isdigit():
                    try:
                        nodes[int(d[4:])] = self._parse_cpulist(open(f"{base}/{d}/cpulist").read())
                    except OSError:
                        continue
        return nodes or {0: list(range(self.logical_cores))}

    @staticmethod
    def _cpu_times():
        try:
            with open("/proc/stat") as f:
                v = [int(x) for x in f.readline().split()[1:]]
            idle = v[3] + (v[4] if len(v) > 4 else 0)
            return sum(v), idle
        except (OSError, ValueError, IndexError):
            return 0, 0

    def update_dynamic_info(self) -> bool:
        """Refreshes available RAM, load average, CPU utilisation since the previous call, clock and
        temperature (when the platform exposes them)."""
        total, idle = self._cpu_times()
        prev = getattr(self, "_prev_times", None)
        self.cpu_utilization = 0.0
        if prev is not None and total > prev[0]:
            self.cpu_utilization = 100.0 * (1.0 - (idle - prev[1]) / (total - prev[0]))
        self._prev_times = (total, idle)
        try:
            self.load_average = os.getloadavg()
        except OSError:
            self.load_average = (0.0, 0.0, 0.0)
        try:
            with open("/proc/meminfo") as f:
                for l in f:
                    if l.startswith("MemAvailable"):
                        self.available_ram_kb = int(l.split()[1])
        except OSError:
            pass
        self.cpu_mhz = 0.0
        try:
            self.cpu_mhz = int(open("/sys/devices/system/cpu/cpu0/cpufreq/scaling_cur_freq").read()) / 1000.0
        except (OSError, ValueError):
            try:
                with open("/proc/cpuinfo") as f:
                    for l in f:
                        if l.startswith("cpu MHz"):
                            self.cpu_mhz = float(l.split(":")[1])
                            break
            except (OSError, ValueError):
                pass
        self.temperature_c = -1.0
        tz = "/sys/class/thermal"
        if os.path.isdir(tz):
            temps = []
            for d in os.listdir(tz):
                try:
                    temps.append(int(open(f"{tz}/{d}/temp").read()) / 1000.0)
                except (OSError, ValueError):
                    continue
            if temps:
                self.temperature_c = max(temps)
        return True

    def to_dict(self) -> dict:
        d = dict(cpu_model=self.cpu_model, sockets=self.sockets, physical_cores=self.physical_cores,
                 logical_cores=self.logical_cores, caches=dict(self.caches), total_ram_kb=self.total_ram_kb,
                 available_ram_kb=self.available_ram_kb, gpus=list(self.gpus),
                 numa={str(k): v for k, v in self.numa_topology().items()})
        for k in ("cpu_utilization", "load_average", "cpu_mhz", "temperature_c"):
            if hasattr(self, k):
                d[k] = getattr(self, k)
        return d

    def to_json(self, indent: int = 2) -> str:
        import json
        return json.dumps(self.to_dict(), indent=indent)


def get_memory_usage_kb() -> int:
    try:
        with open("/proc/self/status") as f:
            for l in f:
                if l.startswith("VmRSS"):
                    return int(l.split()[1])
    except OSError:
        pass
    return 0


def get_shape_hash(shape) -> int:
    return int.from_bytes(hashlib.blake2b(",".join(str(int(s)) for s in shape).encode(), digest_size=8).digest(), "little")


class ThreadAffinity:
    """Pin the calling thread / process to a CPU set (reference pins P/E cores)."""

    @staticmethod
    def pin(cpus) -> bool:
        try:
            os.sched_setaffinity(0, set(int(c) for c in cpus))
            return True
        except (AttributeError, OSError):
            return False

    @staticmethod
    def current() -> set:
        try:
            return set(os.sched_getaffinity(0))
        except AttributeError:
            return set(range(os.cpu_count() or 1))
