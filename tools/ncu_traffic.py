"""DRAM traffic per launch, by kernel, from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv --log-file X` launch list of one step (tools/profile_step.py).

    python tools/ncu_traffic.py launches.csv [kernel-regex] [out.json]

Prints a per-kernel table (launches, time, read + write bytes, GB/s) and, with out.json, writes the aggregate of the kernels
matching the regex (default: gemm_tcgen05) in the form bench.py reads for `roofline.traffic`."""
import collections, csv, json, re, sys

rows = list(csv.reader(open(sys.argv[1])))
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else "gemm_tcgen05")
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]; ix = {n: i for i, n in enumerate(h)}
per = collections.defaultdict(dict)  # launch id -> {metric: value in base units, name}
for r in rows[hi + 1:]:
    if len(r) != len(h):
        continue
    v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
    mult = {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0, "s": 1.0,
            "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u)
    if mult is None:
        raise SystemExit(f"unknown unit {u!r}")
    d = per[r[ix["ID"]]]
    d["name"] = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
    d[r[ix["Metric Name"]]] = v * mult
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in per.values():
    a = agg[d["name"]]
    a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0); a[3] += d.get("dram__bytes_write.sum", 0.0)
print(f"{'ms':>8s} {'n':>5s} {'read MB':>10s} {'write MB':>10s} {'GB/s':>8s}  kernel")
for n, (c, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t * 1e3:8.3f} {c:5d} {rd / 1e6:10.1f} {wr / 1e6:10.1f} {(rd + wr) / max(t, 1e-12) / 1e9:8.0f}  {n[:90]}")
sel = [(n, v) for n, v in agg.items() if pat.search(n)]
c = sum(v[0] for _, v in sel); t = sum(v[1] for _, v in sel); rd = sum(v[2] for _, v in sel); wr = sum(v[3] for _, v in sel)
summary = {"source": f"{sys.argv[1]}: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                     "--clock-control none over one cfg2 step (tools/profile_step.py); serialised, cold-ish caches",
           "kernel_regex": pat.pattern, "launches": c, "time_ms_serialised": t * 1e3, "dram_read_bytes": rd, "dram_write_bytes": wr,
           "mean_dram_bytes_per_launch": (rd + wr) / max(c, 1), "kernels": {n: {"launches": v[0], "ms": v[1] * 1e3,
                                                                               "read": v[2], "write": v[3]} for n, v in sel}}
print(json.dumps({k: v for k, v in summary.items() if k != "kernels"}, indent=1))
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(json.dumps(summary, indent=1))
