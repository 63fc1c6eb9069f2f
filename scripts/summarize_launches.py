"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name, share of the step."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
total = 0.0
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    if unit in ("us", "usecond"): v *= 1e3
    elif unit in ("ms", "msecond"): v *= 1e6
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v; total += v
print("total kernel time %.3f ms over %d launches" % (total / 1e6, sum(a[0] for a in agg.values())))
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%8.3f ms %5.1f%% %5d x %8.1f us  %s" % (t / 1e6, 100 * t / total, n, t / n / 1e3, name[:110]))
