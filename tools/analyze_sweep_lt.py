"""Per-launch analysis of tools/gpu_r2_sweep_lt.sh: for every launch of the step, which arms beat the default arm?

    python tools/analyze_sweep_lt.py [gpurun_out/sweep_lt] [min_gain_us]
"""
import glob
import json
import os
import sys

d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep_lt"
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
arms = {}
for f in sorted(glob.glob(os.path.join(d, "arm*.name"))):
    i = int(os.path.basename(f)[3:-5])
    reps = [json.load(open(p)) for p in sorted(glob.glob(os.path.join(d, "arm%d_*.json" % i)))]
    if reps:
        arms[i] = (open(f).read().strip(), reps)
base_name, base = arms[1]
n = len(base[0])
print("default arm: %d launches, %.1f us (sum of in-situ launch times)" % (n, sum(sum(r[i]["ms"] for i in range(n)) for r in base) / len(base) * 1e3))
for i in sorted(arms):
    name, reps = arms[i]
    if len(reps[0]) != n:
        print("%-28s %d launches (different op list)" % (name, len(reps[0])))
        continue
    print("%-28s %.1f us" % (name, sum(sum(r[k]["ms"] for k in range(n)) for r in reps) / len(reps) * 1e3))
print()
for k in range(n):
    b = [r[k]["ms"] * 1e3 for r in base]
    tb = sum(b) / len(b)
    noise = max(b) - min(b)
    wins = []
    for i in sorted(arms):
        if i == 1:
            continue
        name, reps = arms[i]
        if len(reps[0]) != n:
            continue
        t = [r[k]["ms"] * 1e3 for r in reps]
        if max(t) < min(b) - thr:
            wins.append("%s %.1f (v%d)" % (name, sum(t) / len(t), reps[0][k]["variant"]))
    if wins:
        l = base[0][k]
        print("launch %2d kind %d a_mode %d m %d n %d k %d v%d: default %.1f us (+-%.1f) | %s" % (
            k, l["kind"], l["a_mode"], l["m"], l["n"], l["k"], l["variant"], tb, noise / 2, "; ".join(wins)))
