"""Summarise ncu outputs under gpurun_out/ into profiles/ (tracked).
usage: python scripts/summarize_ncu.py <tag>   (reads gpurun_out/launches.csv, gpurun_out/prof_*.ncu-rep)"""
import collections, csv, glob, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs("profiles", exist_ok=True)
lc = "gpurun_out/launches.csv"
if os.path.exists(lc):
    lines = [l for l in open(lc) if l.startswith('"')]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        k = row["Kernel Name"][:110]
        tot[k][0] += 1; tot[k][1] += float(row["Metric Value"])
    s = sum(v[1] for v in tot.values())
    with open(f"profiles/{tag}_launch_summary.txt", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --timed-only\n")
        f.write(f"# per-kernel totals over the whole run (cold-cache, serialised: compare SHARES); total {s/1e6:.2f} ms, {sum(v[0] for v in tot.values())} launches\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]/1e6:10.3f} ms {100*v[1]/s:6.2f}%  n={v[0]:5d}  avg={v[1]/v[0]/1e3:9.1f} us  {k}\n")
    print(open(f"profiles/{tag}_launch_summary.txt").read()[:1500])
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum"]
def raw_csv(rep):
    """raw page of a report: the CSV made on the GPU box if it is there (large reports do not travel), else ncu -i."""
    pre = rep[:-8] + ".raw.csv"
    if os.path.exists(pre):
        return open(pre).read()
    return subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout


REPS = sorted(set(glob.glob("gpurun_out/prof_*.ncu-rep")) | {f[:-8] + ".ncu-rep" for f in glob.glob("gpurun_out/prof_*.raw.csv")})
for rep in REPS:
    name = os.path.basename(rep)[:-8]
    raw = raw_csv(rep)
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(f"profiles/{tag}_{name}_ncu.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ; one block per captured launch ({name})\n")
        for r in rows[2:]:
            f.write("---\n")
            for w in want:
                if w in idx:
                    f.write(f"{w} = {r[idx[w]][:100]} {units[idx[w]]}\n")
    print("wrote", f"profiles/{tag}_{name}_ncu.txt")

# ---- DRAM traffic per launch of the captured kernels (feeds bench.py's roofline.traffic) ----
import json, re
traffic = {}
def _num(x):
    try: return float(x.replace(",", ""))
    except Exception: return None
for rep in REPS:
    raw = raw_csv(rep)
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3: continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        name = re.sub(r"<.*", "", r[idx["Kernel Name"]]).replace("void ", "").strip()
        rd, wr = _num(r[idx["dram__bytes_read.sum"]]), _num(r[idx["dram__bytes_write.sum"]])
        if rd is None or wr is None: continue
        tot = rd * scale.get(units[idx["dram__bytes_read.sum"]], 1.0) + wr * scale.get(units[idx["dram__bytes_write.sum"]], 1.0)
        d = traffic.setdefault(name, {"dram_bytes": 0.0, "launches_captured": 0, "time_us": 0.0})
        d["dram_bytes"] += tot; d["launches_captured"] += 1
        t = _num(r[idx["gpu__time_duration.sum"]]); tu = units[idx["gpu__time_duration.sum"]]
        d["time_us"] += t * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(tu, 1.0)
for k, d in traffic.items():
    d["dram_bytes_per_launch"] = d["dram_bytes"] / d["launches_captured"]
json.dump(traffic, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print(json.dumps(traffic, indent=1))
