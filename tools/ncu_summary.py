"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), float(r["Metric Value"].replace(",", ""))))
agg = collections.OrderedDict()
for k, ns in rows:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.1f%% |" % (k, n, ns / 1e3, 100 * ns / tot))
print("| TOTAL | %d | %.1f | 100%% |" % (len(rows), tot / 1e3))
