"""Per-launch summary of an `ncu --set full` report (CSV of --page raw): duration, DRAM traffic and
throughput, SM / tensor-pipe utilisation, occupancy, registers.  Usage:
    ncu -i rep.ncu-rep --page raw --csv | python tools/ncu_kernel_summary.py > profiles/xxx.md"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(n for _, n in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        name = name.replace("void ", "").split("(")[0][:60]
        vals = []
        for key, _ in COLS:
            if key in col:
                v, u = r[col[key]], units[col[key]]
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:.4g}"
                except ValueError:
                    pass
                vals.append(f"{v} {u}".strip())
            else:
                vals.append("-")
        print(f"| {name} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
