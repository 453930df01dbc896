#!/usr/bin/env python
"""Top stall sites of one kernel from an ncu report's source page (read here, without a GPU).

    python tools/ncu_source_top.py <report.ncu-rep> <kernel-index> [top-N] [--cuda]
"""
import csv
import io
import subprocess
import sys

rep, kidx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 40
mode = "cuda" if "--cuda" in sys.argv else "sass"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", mode],
                     capture_output=True, text=True).stdout
sections, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
sec = sections[kidx]
hdr = sec["hdr"]
si = hdr.index("# Samples")
src = hdr.index("Source")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in sec["rows"])
print(sec["name"][:120], "total samples", tot)
agg = {}
for i, h in stall_cols:
    agg[h] = sum(int(r[i] or 0) for r in sec["rows"])
print("stall totals:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(tot, 1)) for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
rows = sorted(sec["rows"], key=lambda r: -int(r[si] or 0))[:top]
for r in rows:
    st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:3]
    print("%6d %5.1f%%  %-70s %s" % (int(r[si] or 0), 100.0 * int(r[si] or 0) / max(tot, 1), r[src].strip()[:70],
                                     " ".join("%s:%d" % (h, v) for v, h in st if v)))
