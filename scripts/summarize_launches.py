"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals + ordered list."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); seq = []; tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:64]
    agg[name][0] += 1; agg[name][1] += v; tot += v; seq.append((name, v, row.get("Grid Size")))
print(f"{len(seq)} launches, {tot/1e3:.3f} ms total")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"{k:66s} {n:4d} {t:9.1f} us {100*t/tot:5.1f}%")
if len(sys.argv) > 3:
    for i, s in enumerate(seq):
        if any(x in s[0] for x in sys.argv[3].split(",")): print(i, s[0], f"{s[1]:.1f}", s[2])
