"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (and grid)."""
import csv, collections, re, sys
path = sys.argv[1]
by_grid = len(sys.argv) > 2
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void rb::", "").replace("rb::", "")
    key = (name, row["Grid Size"]) if by_grid else name
    agg[key][0] += 1
    agg[key][1] += v
    tot += v
print(f"total {tot/1e3:.2f} ms over {sum(n for n, _ in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{t:10.1f} us {100*t/tot:5.1f}%  n={n:5d} avg={t/n:8.1f}  {k}")
