"""Aggregate an ncu per-launch CSV (gpu__time_duration.sum [+ dram bytes]) by kernel name and by (kernel, grid).
usage: python scripts/summarize_launches.py gpurun_out/train_launches.csv [out.txt]"""
import collections, csv, sys
rows = list(csv.DictReader([l for l in open(sys.argv[1]) if l.startswith('"')]))
per = collections.OrderedDict()
for r in rows:
    key = r["ID"]
    d = per.setdefault(key, {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
tot = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in per.values():
    t = tot[d["name"][:90]]
    t[0] += 1; t[1] += d.get("gpu__time_duration.sum", 0.0); t[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
s = sum(v[1] for v in tot.values())
out = [f"# total {s / 1e3:.3f} ms over {len(per)} launches (ncu per-launch times: cold cache, serialised)"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{v[1] / 1e3:9.3f} ms {100 * v[1] / s:6.2f}%  n={v[0]:5d}  avg={v[1] / v[0]:8.1f} us  dram={v[2] / 1e6:9.1f} MB  {v[2] / max(v[1], 1e-9) / 1e3:7.1f} GB/s  {k}")
out.append("")
out.append("# per (kernel, grid): launches, total us, avg us, dram MB per launch, GB/s")
tg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in per.values():
    t = tg[(d["name"][:60], d["grid"])]
    t[0] += 1; t[1] += d.get("gpu__time_duration.sum", 0.0); t[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
for k, v in sorted(tg.items(), key=lambda kv: -kv[1][1])[:70]:
    out.append(f"{v[1]:9.1f} us  n={v[0]:4d}  avg={v[1] / v[0]:8.1f} us  {v[2] / v[0] / 1e6:8.1f} MB  {v[2] / max(v[1], 1e-9) / 1e3:7.1f} GB/s  {k[0]} grid={k[1]}")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
