#!/usr/bin/env python
"""Summarise ncu captures (read here, without a GPU) into markdown for profiles/.

    python tools/ncu_summary.py --rep gpurun_out/prof.ncu-rep [--rep ...] --launches gpurun_out/launches.csv --out profiles/rNN_x.md
"""
import argparse
import csv
import io
import subprocess
from collections import OrderedDict

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor hmma % (active)"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor inst %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX %peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    return hdr, units, body


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rep", action="append", default=[])
    ap.add_argument("--launches", action="append", default=[])
    ap.add_argument("--out", required=True)
    ap.add_argument("--title", default="ncu summary")
    ap.add_argument("--note", action="append", default=[])
    a = ap.parse_args()
    md = ["# %s" % a.title, ""]
    for n in a.note:
        md += [n, ""]
    for rep in a.rep:
        hdr, units, body = raw_page(rep)
        md += ["## `%s` (ncu --set full --clock-control none)" % rep, ""]
        col = {h: i for i, h in enumerate(hdr)}
        names = [(k, lab) for k, lab in KEYS if k in col]
        md.append("| # | kernel | " + " | ".join(lab for _, lab in names) + " |")
        md.append("|---|---|" + "---|" * len(names))
        for j, row in enumerate(body):
            kn = row[col["Kernel Name"]].replace("ConvGemmParams", "").replace("void ", "")
            vals = []
            for k, _ in names:
                v, u = row[col[k]], units[col[k]]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                vals.append("%s %s" % (v, u) if u and u != "%" else v)
            md.append("| %d | `%s` | " % (j, kn) + " | ".join(vals) + " |")
        md.append("")
    for launches in a.launches:
            allrows = [r for r in csv.reader(open(launches)) if len(r) > 10]
            hdr = next(r for r in allrows if r[0] == "ID")
            ik, ig = hdr.index("Kernel Name"), hdr.index("Grid Size")
            rows = [r for r in allrows if r[0].isdigit()]
            tot = sum(float(r[-1]) for r in rows)
            agg = OrderedDict()
            for r in rows:
                k = r[ik].split("(")[0].replace("void ", "")
                agg.setdefault(k, [0, 0.0])
                agg[k][0] += 1
                agg[k][1] += float(r[-1])
            md += ["## launch list `%s` (gpu__time_duration.sum, one step, cold-cache/serialised: compare shares)" % launches, "",
                   "| kernel | launches | total us | share of step |", "|---|---|---|---|"]
            for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                md.append("| `%s` | %d | %.1f | %.1f%% |" % (k, c, t / 1e3, 100 * t / tot))
            md += ["", "total %d launches, %.1f us" % (len(rows), tot / 1e3), "", "<details><summary>every launch</summary>", "",
                   "| id | kernel | grid | ns |", "|---|---|---|---|"]
            for r in rows:
                md.append("| %s | `%s` | %s | %s |" % (r[0], r[ik].split("(")[0].replace("void ", ""), r[ig], r[-1]))
            md += ["", "</details>", ""]
    open(a.out, "w").write("\n".join(md))
    print("wrote", a.out)


if __name__ == "__main__":
    main()
