#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python tools/summarise_launches.py launches.csv [title]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hdr_i]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows[hdr_i + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}[r[mu]]
        name = r[kn].split("(")[0][-48:]
        tot[name] += v * scale
        cnt[name] += 1
    total = sum(tot.values())
    print(title)
    print(f"launches {sum(cnt.values())}, summed kernel time {total:.2f} ms (cold-cache, serialised)")
    for name in sorted(tot, key=tot.get, reverse=True):
        print(f"{name:48s} n={cnt[name]:4d} total={tot[name]:9.2f} ms avg={tot[name] / cnt[name] * 1e3:9.1f} us share={100 * tot[name] / total:5.1f}%")


if __name__ == "__main__":
    main()
