#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum,dram__bytes.sum --csv` log into a per-kernel table (markdown):

  python tools/launch_table.py gpurun_out/launches.csv [--skip N] [--title "..."] > profiles/launches_<what>.md
"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    title = sys.argv[sys.argv.index("--title") + 1] if "--title" in sys.argv else path
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    per_id = OrderedDict()
    for r in rows:
        k = r["ID"]
        d = per_id.setdefault(k, {"name": r["Kernel Name"], "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "")
        if r["Metric Name"].startswith("gpu__time_duration"):
            d["us"] = v / 1e3 if unit in ("ns", "nsecond") else v * (1e3 if unit == "msecond" else 1.0)
        elif r["Metric Name"].startswith("dram__bytes"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d["dram"] = d.get("dram", 0) + v * mult
    launches = list(per_id.values())[skip:]
    agg = OrderedDict()
    for l in launches:
        name = l["name"].split("(")[0][:90]
        a = agg.setdefault(name, {"n": 0, "us": 0.0, "dram": 0.0})
        a["n"] += 1
        a["us"] += l.get("us", 0.0)
        a["dram"] += l.get("dram", 0.0)
    total = sum(a["us"] for a in agg.values())
    print(f"# {title}\n")
    print(f"{len(launches)} launches, {total / 1e3:.3f} ms of kernel time (per-kernel durations measured by ncu; not a bench number)\n")
    print("| total us | share | launches | avg us | DRAM MB | GB/s | kernel |\n|---|---|---|---|---|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        bw = a["dram"] / (a["us"] * 1e-6) / 1e9 if a["us"] > 0 and a["dram"] else 0
        print(f"| {a['us']:.1f} | {100 * a['us'] / max(total, 1e-9):.1f}% | {a['n']} | {a['us'] / a['n']:.1f} | {a['dram'] / 1e6:.1f} | {bw:.0f} | `{name}` |")


if __name__ == "__main__":
    main()
