"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python scripts/launch_table.py launches.csv [steps]"""
import collections, csv, re, sys
path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
with open(path) as fh:
    lines = [l for l in fh if not l.startswith("==")]
agg = collections.OrderedDict()
n = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ms = v / 1e6 if unit == "ns" else (v / 1e3 if unit == "us" else v)
    name = re.sub(r"^void ", "", row["Kernel Name"]).replace("<unnamed>::", "")
    name = re.sub(r"\(.*", "", name)[:60]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
    n += 1
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | share | ms per launch | ms per step |\n|---|---:|---:|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    print("| `%s` | %d | %.3f | %.2f %% | %.3f | %.3f |" % (k, a[0], a[1], 100 * a[1] / tot, a[1] / a[0], a[1] / steps))
print("\n%d launches, %.1f ms in total (%d steps: %.2f ms per step, serialised and cold-cache: compare SHARES)" % (n, tot, steps, tot / steps))
