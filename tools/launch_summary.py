"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]; ix = {n: i for i, n in enumerate(h)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) != len(h) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v for _, v in agg.values())
print(f"total {tot / 1e3:.3f} ms over {sum(c for c, _ in agg.values())} launches")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v / 1e3:8.3f} ms {c:5d} x {v / c:8.1f} us  {n[:100]}")
