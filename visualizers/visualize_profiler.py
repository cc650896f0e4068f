#!/usr/bin/env python
"""Gantt chart of a merged profiler log (reference: visualizers/visualize_profiler.py:8-70).
Parses ``logs/profiler.log`` as written by ``Profiler.log`` / ``Coordinator.print_profiling`` and
draws one lane per source with COMPUTE and COMMUNICATION spans.  Falls back to a text summary when
matplotlib is not installed.

  python visualizers/visualize_profiler.py logs/profiler.log [out.png]
"""
import re
import sys
from collections import defaultdict

LINE = re.compile(r"\[(?P<src>[^\]]*)\]\s+(?P<type>COMPUTE|COMMUNICATION|OTHER)\s+(?P<name>.*?)\s+start=(?P<s>[-\d.]+)ms\s+end=(?P<e>[-\d.]+)ms")


def parse(path):
    ev = []
    for line in open(path):
        m = LINE.search(line)
        if m:
            ev.append((m["src"] or "main", m["type"], m["name"], float(m["s"]), float(m["e"])))
    return ev


def summarize(ev):
    tot = defaultdict(float)
    for src, typ, _, s, e in ev:
        tot[(src, typ)] += e - s
    for (src, typ), ms in sorted(tot.items()):
        print(f"{src:24s} {typ:14s} {ms:10.3f} ms")


if __name__ == "__main__":
    path = sys.argv[1] if len(sys.argv) > 1 else "logs/profiler.log"
    ev = parse(path)
    print(f"{len(ev)} events from {path}")
    summarize(ev)
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:  # noqa: BLE001
        sys.exit(0)
    lanes = sorted({e[0] for e in ev})
    fig, ax = plt.subplots(figsize=(14, 1 + 0.6 * len(lanes)))
    colors = {"COMPUTE": "tab:blue", "COMMUNICATION": "tab:orange", "OTHER": "tab:gray"}
    for src, typ, name, s, e in ev:
        ax.barh(lanes.index(src), e - s, left=s, color=colors.get(typ, "tab:gray"), edgecolor="none")
    ax.set_yticks(range(len(lanes)))
    ax.set_yticklabels(lanes)
    ax.set_xlabel("ms since profiling start")
    out = sys.argv[2] if len(sys.argv) > 2 else "profiler.png"
    fig.savefig(out, dpi=120, bbox_inches="tight")
    print("wrote", out)
