#!/usr/bin/env python
"""Markdown table of the judged metrics from `ncu -i X.ncu-rep --page raw --csv` exports.
Usage: python tools/ncu_summary.py a.csv [b.csv ...]"""
import csv
import sys

COLS = [
    ("time us", "gpu__time_duration.sum", 1.0),
    ("DRAM read MB", "dram__bytes_read.sum", None),
    ("DRAM write MB", "dram__bytes_write.sum", None),
    ("DRAM thr %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("L2 thr %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("tensor pipe active %", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
    ("ALU %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1.0),
    ("FMA %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1.0),
    ("XU %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0),
    ("warp instr M", "smsp__inst_executed.sum", 1e-6),
    ("regs", "launch__registers_per_thread", 1.0),
    ("dyn smem KB", "launch__shared_mem_per_block_dynamic", 1.0),
]


def to_mb(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1.0)


def main():
    print("| kernel | grid | " + " | ".join(c[0] for c in COLS) + " |")
    print("|---|---|" + "---:|" * len(COLS))
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        h, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[h.index("Kernel Name")]
            grid = r[h.index("Grid Size")] if "Grid Size" in h else ""
            cells = []
            for label, key, scale in COLS:
                if key not in h:
                    cells.append("n/a")
                    continue
                i = h.index(key)
                if scale is None:
                    cells.append("%.1f" % to_mb(r[i], units[i]))
                else:
                    v = float(r[i].replace(",", ""))
                    if label == "time us" and units[i] == "ms":
                        v *= 1000.0
                    if label == "time us" and units[i] == "ns":
                        v /= 1000.0
                    cells.append("%.1f" % (v * scale))
            print("| `%s` | %s | %s |" % (name[:60], grid, " | ".join(cells)))


if __name__ == "__main__":
    main()
