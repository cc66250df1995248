"""Turn the raw outputs of tools/gpu_profile_r02.sh (gpurun_out/r02_*) into the small tracked summaries under profiles/."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
OUT = os.path.join(ROOT, "profiles")
tag = "r02"

# 1. launch list -> per-kernel totals and shares of the LAST solve
rows = [r for r in csv.reader(open(os.path.join(GO, f"{tag}_launches.csv"))) if len(r) > 5]
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
timed = [r for r in rows[1:] if r[im] == "gpu__time_duration.sum"]
resolve = [i for i, r in enumerate(timed) if "walkResolveKernel" in r[ik]]
# the bench runs 3 warm-up solves, 1 timed solve, then the e2e leg (1 warm + 1 timed): keep the 4th solve (the timed one of the value leg)
start = (resolve[2] + 1) if len(resolve) >= 4 else 0
stop = (resolve[3] + 1) if len(resolve) >= 4 else len(timed)
agg = collections.OrderedDict()
for r in timed[start:stop]:
    name = r[ik].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", ""))
unit = rows[1][hdr.index("Metric Unit")]
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
tot = sum(a[1] for a in agg.values())
with open(os.path.join(OUT, f"{tag}_launches_summary.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-verify --no-extras\n")
    f.write("# the timed solve of the value leg (kernels between the 3rd and the 4th walkResolveKernel); times under ncu are cold-cache and serialised: compare SHARES\n")
    f.write(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k[:70]:70s} {n:8d} {t * scale:12.1f} {t * scale / n:10.2f} {100 * t / tot:6.1f}%\n")
print(open(os.path.join(OUT, f"{tag}_launches_summary.txt")).read())

# 2. full captures -> key metrics
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg']
what = {"r02_prof_res": "bench.py --T 400 (BigRoom 1024^2, 4 sources; ONE launch = one source, 100 passes x 4 steps)",
        "r02_prof_ws2": "tools/gpu_time_one.py HugeRoom 2048 400 2 0 (config 4's grid: 2048^2, two sources; one launch = 100 generations x 4 steps)",
        "r02_prof_encode": "bench.py --T 400",
        "r02_prof_encode_huge": "tools/gpu_time_one.py HugeRoom 2048 4000 2 0 (an all-onset scene at full length: two 2048^2 sources x 4000 steps, 134 GB of history)"}
with open(os.path.join(OUT, f"{tag}_ncu_kernels.txt"), "w") as f:
    for rep in ("r02_prof_res", "r02_prof_ws2", "r02_prof_encode", "r02_prof_encode_huge"):
        path = os.path.join(GO, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rr = list(csv.reader(out.splitlines()))
        h, u = rr[0], rr[1]
        for r in rr[2:]:
            f.write(f"--- {r[h.index('Kernel Name')][:110]}   [{rep}.ncu-rep: ncu --set full --clock-control none --import-source on, {what[rep]}]\n")
            for w in want:
                if w in h:
                    f.write(f"   {w:80s} {r[h.index(w)]:>18s} {u[h.index(w)]}\n")
            for i, name in enumerate(h):
                if 'issue_stalled' in name and name.endswith('per_issue_active.ratio'):
                    try:
                        v = float(r[i])
                    except ValueError:
                        continue
                    if v >= 0.2:
                        f.write(f"   stall {name.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:8.2f} warps/issue\n")
print(open(os.path.join(OUT, f"{tag}_ncu_kernels.txt")).read()[:4000])
for rep, dst in (("r02_prof_ws2", "r02_ws2_stalls_by_line.txt"), ("r02_prof_res", "r02_resident_stalls_by_line.txt")):
    path = os.path.join(GO, rep + ".ncu-rep")
    if os.path.exists(path):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_stalls_by_line.py"), path], capture_output=True, text=True).stdout
        open(os.path.join(OUT, dst), "w").write(f"# python tools/ncu_stalls_by_line.py gpurun_out/{rep}.ncu-rep   (warp-stall samples per CUDA source line)\n" + out)
for name in ("r02_bench_default.json", "r02_bench_reference.json"):
    src = os.path.join(GO, name)
    if os.path.exists(src):
        open(os.path.join(OUT, name), "w").write(open(src).read())
