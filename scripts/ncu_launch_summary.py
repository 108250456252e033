"""Per-kernel totals of the LAST of `reps` identical passes in an ncu launch list (gpu__time_duration.sum)."""
import collections
import csv
import sys

path, reps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = list(csv.DictReader([l for l in open(path) if not l.startswith("==")]))
rows = rows[len(rows) - len(rows) // reps:]
agg, tot = collections.OrderedDict(), 0.0
for row in rows:
    name = row["Kernel Name"].replace("unnamed>::", "")[:64]
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    a = agg.setdefault(name, [0.0, 0])
    a[0] += v
    a[1] += 1
    tot += v
for k, (v, c) in agg.items():
    print(f"{v:10.1f} us  x{c:<3d} {100 * v / tot:5.1f} %  {k}")
print(f"{tot:10.1f} us  total, {len(rows)} launches")
