#!/usr/bin/env python
"""Summarise an Nsight Compute report into a small table for profiles/ (run here, no GPU needed):

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/ncu_summary_<what>.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "dur_us"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_%"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform_inst"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall_long_sb"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall_barrier"),
    ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall_lg_throttle"),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "stall_mio"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {}
    for key, short in METRICS:
        for i, h in enumerate(hdr):
            if h == key:
                idx[short] = i
                break
    ki, gi, bi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    cols = [s for _, s in METRICS if s in idx]
    print("| kernel | grid | block | " + " | ".join(f"{c} ({units[idx[c]]})" if units[idx[c]] else c for c in cols) + " |")
    print("|---|---|---|" + "---|" * len(cols))
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("tnn::", "")[:60]
        vals = []
        for c in cols:
            v = r[idx[c]]
            try:
                vals.append(f"{float(v.replace(',', '')):.4g}")
            except ValueError:
                vals.append(v)
        print(f"| {name} | {r[gi]} | {r[bi]} | " + " | ".join(vals) + " |")
    # ---- top warp-stall reasons per distinct kernel (cycles a warp waits per issued instruction)
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warp") and "issue_stalled_" in h and h.endswith(".ratio")]
    if stall_cols:
        print("\nTop stall reasons per kernel (warp cycles per issued instruction; first instance of each kernel):\n")
        print("| kernel | reasons |")
        print("|---|---|")
        seen = set()
        for r in rows[2:]:
            name = r[ki].split("(")[0].replace("void ", "").replace("tnn::", "")[:60]
            if name in seen:
                continue
            seen.add(name)
            vals = []
            for i in stall_cols:
                try:
                    vals.append((float(r[i].replace(",", "")), hdr[i].split("issue_stalled_")[1].split(".")[0].replace("_per_issue_active", "")))
                except ValueError:
                    pass
            vals.sort(reverse=True)
            print(f"| {name} | " + ", ".join(f"{b} {a:.2f}" for a, b in vals[:5]) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
