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
    ours = {k: v for k, v in agg.items() if k.startswith("srk::")}
    other = {k: v for k, v in agg.items() if not k.startswith("srk::")}
    tot = sum(a[1] for a in ours.values())
    print("Kernels of this repository (`srk::`), share = share of the summed kernel time of the engine:\n")
    print("| kernel | launches | total ms | avg ms | max ms | share |")
    print("|---|---:|---:|---:|---:|---:|")
    for name, a in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.4f | %.4f | %.1f %% |" % (name[5:], a[0], a[1], a[1] / a[0], a[2], 100 * a[1] / tot))
    print("| **all srk kernels** | %d | %.3f | | | |" % (sum(a[0] for a in ours.values()), tot))
    if other:
        print("\nOther kernels in the same process (bench.py's own cuBLAS DGEMM 8192^3 peak measurement and torch helpers; not part of a step):\n")
        print("| kernel | launches | total ms | avg ms |")
        print("|---|---:|---:|---:|")
        for name, a in sorted(other.items(), key=lambda kv: -kv[1][1]):
            print("| `%s` | %d | %.3f | %.4f |" % (name[:80], a[0], a[1], a[1] / a[0]))


if __name__ == "__main__":
    main()
