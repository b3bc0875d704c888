#!/usr/bin/env python
"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).

  python tools/summarize_launches.py gpurun_out/launches_X.csv > profiles/r01_launches_X.md
"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    out = []
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
        out.append((r[ki].split("(")[0].replace("void ", ""), v))
    return out


def main():
    path = sys.argv[1]
    launches = load(path)
    agg = collections.OrderedDict()
    for name, ms in launches:
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1; a[1] += ms; a[2] = max(a[2], ms)
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | avg ms | max ms | share |")
    print("|---|---:|---:|---:|---:|---:|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.4f | %.4f | %.1f %% |" % (name, a[0], a[1], a[1] / a[0], a[2], 100 * a[1] / tot))
    print("| **all** | %d | %.3f | | | |" % (len(launches), tot))


if __name__ == "__main__":
    main()
