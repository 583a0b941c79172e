#!/usr/bin/env python
"""Compact per-kernel table from an `ncu --page raw --csv` export: python profiles/summarize_ncu.py file_raw.csv ..."""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB rd"), ("dram__bytes_write.sum", "MB wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def to_num(v, unit, want):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3}
    if want.startswith("MB") or want == "us":
        x *= scale.get(unit, 1.0)
    return round(x, 2)


for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"== {path}")
    print(f"{'kernel':44s} " + " ".join(f"{n:>8s}" for _, n in KEYS))
    for r in rows[2:]:
        vals = []
        for key, name in KEYS:
            if key in hdr:
                i = hdr.index(key)
                vals.append(to_num(r[i], units[i], name))
            else:
                vals.append("-")
        print(f"{r[ki][:44]:44s} " + " ".join(f"{str(v):>8s}" for v in vals))
