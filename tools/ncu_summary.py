#!/usr/bin/env python
"""Condense an `ncu -i X.ncu-rep --page raw --csv` dump into the handful of per-launch metrics the
roofline discussion uses.   python tools/ncu_summary.py raw.csv > profiles/NAME.md"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__t_sectors_op_read.sum", "l2_rd_sect"),
    ("lts__t_sectors_op_write.sum", "l2_wr_sect"),
    ("lts__t_sector_hit_rate.pct", "l2_hit%"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
    ("smsp__inst_executed.sum", "inst"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    have = [(hdr.index(c), n, units[hdr.index(c)]) for c, n in COLS if c in hdr]
    ki, gi, bi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    print("| # | kernel | grid | block | " + " | ".join(f"{n} [{u}]" if u else n for _, n, u in have) + " |")
    print("|" + "---|" * (4 + len(have)))
    for k, r in enumerate(rows[2:]):
        name = r[ki].split("(")[0].replace("void ", "")
        vals = []
        for i, _, _ in have:
            try:
                v = float(r[i].replace(",", ""))
                vals.append(f"{v:.4g}")
            except ValueError:
                vals.append(r[i])
        print(f"| {k} | {name} | {r[gi]} | {r[bi]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
