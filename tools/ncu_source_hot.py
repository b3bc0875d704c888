#!/usr/bin/env python
"""Hot source lines (warp-stall samples) of one kernel in an ncu report:  tools/ncu_source_hot.py rep.ncu-rep <kernel-substring> [top] [nth-match]"""
import csv, io, subprocess, sys
rep, name = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
blocks, cur, fpath = [], None, ""
for row in csv.reader(io.StringIO(raw)):
    if not row: continue
    if row[0] == "File Path":
        fpath = row[1]; continue
    if row[0] == "Function Name":
        cur = {"fn": row[1], "rows": [], "hdr": None, "file": fpath}; blocks.append(cur); continue
    if cur is None: continue
    if row[0] == "Line No": cur["hdr"] = row; continue
    if cur["hdr"] is not None: cur["rows"].append(row)
fpath = ""
m = [b for b in blocks if name in b["fn"]]
# one block per (file, function): merge them, tagging every line with its file
h = m[0]["hdr"]
b = {"fn": m[0]["fn"], "rows": []}
for blk in m:
    tag = blk["file"].split("/")[-1]
    for r in blk["rows"]:
        if r and r[0]: r = [tag + ":" + r[0]] + r[1:]
        b["rows"].append(r)
si = h.index("# Samples") - len(h)          # index from the end: a source line with quotes may split the Source column
stall_cols = [i - len(h) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
def num(x):
    try: return int(float(x or 0))
    except ValueError: return 0
rows = [r for r in b["rows"] if r[0] and len(r) >= len(h)]   # source lines only (SASS rows have an empty line number)
tot = sum(num(r[si]) for r in rows)
print(b["fn"][:100], "total samples", tot)
rows = [r for r in rows if num(r[si]) > 0]
rows.sort(key=lambda r: -num(r[si]))
for r in rows[:top]:
    st = sorted(((num(r[i]), h[i][6:]) for i in stall_cols), reverse=True)[:3]
    print("%-22s %6d %5.1f%%  %-90s %s" % (r[0], num(r[si]), 100.0 * num(r[si]) / max(tot, 1), r[1].strip()[:90], " ".join("%s=%d" % (n, v) for v, n in st if v)))
