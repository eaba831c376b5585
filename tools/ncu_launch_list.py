"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel counts, totals and shares.
usage: python tools/ncu_launch_list.py launches.csv "<command that was profiled>" [first_launch last_launch] > profiles/<name>.md"""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else len(rows)
rows = rows[lo:hi]
agg = {}
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("npe::", "")
    d = agg.setdefault(name, [0, 0.0, r[8]])
    d[0] += 1; d[1] += float(r[-1]) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"# ncu launch list (gpu__time_duration.sum), `{sys.argv[2]}`, launches {lo}-{hi - 1}\n")
print("Per-launch times are cold-cache and serialised under the profiler (no programmatic-dependent-launch overlap): compare SHARES, not absolutes.\n")
print("| kernel | grid | launches | total us | mean us | share |\n|---|---|---|---|---|---|")
for k, (n, t, g) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {g} | {n} | {t:.1f} | {t / n:.2f} | {100 * t / tot:.1f}% |")
