#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time and
share of the captured region.  Usage: ncu_launch_summary.py launches.csv [first_id last_id] > summary.txt"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void\s+", "", name)
    m = re.match(r"(?:dv::)?([A-Za-z0-9_:]+)", name)
    base = m.group(1) if m else name[:60]
    t = re.search(r"<(.*)>\(", name)
    if base.startswith("at::") or "at::native" in name:
        return "torch:" + (re.search(r"(\w+_kernel\w*|\w+Functor\w*)", name) or [name[:40]])[0]
    return base + ("<" + t.group(1)[:48] + ">" if t else "")


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) >= 15 and r[0].isdigit()]
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        i = int(r[0])
        if i < lo or i > hi or r[12] != "gpu__time_duration.sum":
            continue
        v = float(r[14].replace(",", ""))
        unit = r[13]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        k = short(r[4])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    print(f"# {path}: launches {lo}..{hi if hi < 1 << 60 else 'end'}; {sum(a[0] for a in agg.values())} launches, "
          f"{total / 1e3:.3f} ms summed (ncu: serialised, cold-cache per launch)")
    print(f"{'share':>7} {'ms':>9} {'launches':>8}  kernel")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * us / total:6.2f}% {us / 1e3:9.3f} {n:8d}  {k}")


if __name__ == "__main__":
    main()
