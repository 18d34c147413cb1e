#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report: ncu -i X.ncu-rep --page raw --csv | this."""
import csv
import re
import sys

KEEP = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%peak"),
    ("lts__t_sector_hit_rate.pct", "l2_hit%"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
]


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(n for _, n in KEEP) + " |")
    print("|---|" + "---|" * len(KEEP))
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
        name = re.sub(r"^void |sb::|cub::", "", name)[:44]
        cells = []
        for m, _ in KEEP:
            if m in idx:
                v, u = r[idx[m]], units[idx[m]]
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            else:
                cells.append("-")
        print(f"| {name} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
