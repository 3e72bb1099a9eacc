#!/usr/bin/env python
"""Dump the roofline-relevant metrics of every launch in an .ncu-rep (read with `ncu -i ... --page raw --csv`).
Usage: ncu_report_summary.py report.ncu-rep [...] > summary.txt"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("lts__t_bytes.sum", "l2_bytes"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print(f"## {path}")
        for r in rows[2:]:
            name = r[col["Kernel Name"]].split("(")[0]
            parts = [name]
            for m, short in WANT:
                if m in col:
                    parts.append(f"{short}={r[col[m]]}{units[col[m]]}")
            print("  " + "  ".join(parts))


if __name__ == "__main__":
    main()
