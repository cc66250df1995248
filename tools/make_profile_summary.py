"""Turn the raw outputs of tools/gpu_profile.sh (gpurun_out/) into the small tracked summaries under profiles/."""
import csv, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)

# 1. launch list -> per-kernel totals and shares
rows = [r for r in csv.reader(open(os.path.join(GO, "launches.csv"))) if len(r) > 5]
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
timed = [r for r in rows[1:] if r[im] == "gpu__time_duration.sum"]
# the last two solves of the run: cut at the third-from-last encodeResponseKernel's end, i.e. keep everything after the
# launch that follows it (the step kernels of a solve precede its analyzer kernels)
enc = [i for i, r in enumerate(timed) if "encodeResponseKernel" in r[ik]]
resolve = [i for i, r in enumerate(timed) if "walkResolveKernel" in r[ik] or "listenerDirectionKernel" in r[ik]]
start = (resolve[-3] + 1) if len(resolve) >= 3 else 0
for r in timed[start:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = r[ik].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", ""))
unit = rows[1][hdr.index("Metric Unit")]
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)       # -> microseconds
tot = sum(a[1] for a in agg.values())
with open(os.path.join(OUT, f"{tag}_launches_summary.txt"), "w") as f:
    f.write("# PVC_NO_GRAPHS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 python bench.py --steps 1 --warmup 3 --no-cpu-baseline\n")
    f.write("# the last two solves of that run (1 solve = 4 step-kernel launches of <= 256 generations x 4 time steps + 8 analyzer kernels); cold-cache, serialised: compare SHARES\n")
    f.write(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>9s} {'share':>7s}\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k[:60]:60s} {n:8d} {t*scale:12.1f} {t*scale/n:9.2f} {100*t/tot:6.1f}%\n")
print(open(os.path.join(OUT, f"{tag}_launches_summary.txt")).read())

# 2. full captures -> key metrics
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg']
traffic = None
with open(os.path.join(OUT, f"{tag}_ncu_kernels.txt"), "w") as f:
    for rep in ("prof_fused", "prof_encode", "prof_walk"):
        path = os.path.join(GO, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rr = list(csv.reader(out.splitlines()))
        h, u = rr[0], rr[1]
        for r in rr[2:]:
            f.write(f"--- {r[h.index('Kernel Name')][:100]}   [{rep}.ncu-rep: ncu --set full --clock-control none, bench.py --T 400]\n")
            for w in want:
                if w in h:
                    f.write(f"   {w:70s} {r[h.index(w)]:>18s} {u[h.index(w)]}\n")
            for i, name in enumerate(h):
                if 'issue_stalled' in name and name.endswith('per_issue_active.ratio'):
                    try:
                        v = float(r[i])
                    except ValueError:
                        continue
                    if v >= 0.2:
                        f.write(f"   stall {name.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):30s} {v:8.2f} warps/issue\n")
            if rep == "prof_fused" and traffic is None:
                def val(n):
                    x = float(r[h.index(n)].replace(",", "")); un = u[h.index(n)]
                    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(un, 1)
                traffic = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
print(open(os.path.join(OUT, f"{tag}_ncu_kernels.txt")).read()[:3000])
if traffic:
    json.dump({"kernel": "pvc::ws2::stepKernel<14,4,1,false,true>", "dram_bytes_per_launch": traffic, "generations_in_that_launch": 100,
               "dram_bytes_per_generation": traffic / 100.0, "source": f"profiles/{tag}_ncu_kernels.txt (ncu --set full, cold L2)",
               "note": "captured with --T 400: ONE launch = 100 generations x 4 steps, 4 sources, 1024x1024", "algorithmic_bytes_per_launch": 28 * 1024 * 1024 * 4 * 400}, open(os.path.join(OUT, "fused_step_traffic.json"), "w"), indent=1)
for name in ("bench_default.json", "timeline.txt"):
    src = os.path.join(GO, name)
    if os.path.exists(src):
        open(os.path.join(OUT, f"{tag}_{name}"), "w").write(open(src).read())
