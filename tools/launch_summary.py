"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/launch_summary.py launches.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[h]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    v = float(r[vi].replace(",", ""))
    ns = v * 1e3 if r[ui] == "us" else (v * 1e6 if r[ui] == "ms" else v)
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(a[1] for a in agg.values())
print("# %d launches, %.1f us of kernel time in total (cold-cache, serialised replays: compare shares, not absolutes)" % (sum(a[0] for a in agg.values()), tot / 1e3))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-78s n=%3d  total %10.1f us  avg %9.1f us  share %5.1f%%" % (k[:78], a[0], a[1] / 1e3, a[1] / a[0] / 1e3, 100 * a[1] / tot))
