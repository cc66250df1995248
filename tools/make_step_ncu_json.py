"""profiles/step_kernel_ncu.json from a full ncu capture of the step kernel (one launch): DRAM bytes and issue-slot utilisation per
cell-update, which bench.py scales to the launches of its own run (roofline.traffic / physical_frac / issue_frac).

  python tools/make_step_ncu_json.py gpurun_out/r02_prof_step.ncu-rep --variant 65 --n 1024 --T 400 --sources 1 [--out profiles/step_kernel_ncu.json]
(n, T, sources: what the captured launch solved: cell-updates = n*n*T*sources)"""
import argparse
import csv
import json
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("--variant", type=int, required=True)
ap.add_argument("--n", type=int, required=True)
ap.add_argument("--T", type=int, required=True)
ap.add_argument("--sources", type=int, default=1)
ap.add_argument("--out", default="profiles/step_kernel_ncu.json")
ap.add_argument("--command", default="")
args = ap.parse_args()

out = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(name):
    v, u = m[name]
    x = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    return x * scale


cu = args.n * args.n * args.T * args.sources
dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
doc = {
    "variant": args.variant, "grid_n": args.n, "kernel": m["Kernel Name"][0] if "Kernel Name" in m else "",
    "captured_launch": {"cell_updates": cu, "time_steps": args.T, "sources": args.sources, "duration_s_under_ncu": num("gpu__time_duration.sum"),
                        "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum")},
    "dram_bytes_per_cell_update": dram / cu,
    "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
    "warps_active_pct": float(m["sm__warps_active.avg.pct_of_peak_sustained_active"][0]),
    "thread_instructions_per_cell_update": 32.0 * float(m["smsp__inst_executed.sum"][0].replace(",", "")) / cu,
    "l2_hit_rate_pct": float(m["lts__t_sector_hit_rate.pct"][0]),
    "registers_per_thread": int(float(m["launch__registers_per_thread"][0])),
    "source": f"{args.report.split('/')[-1]}: ncu --set full --clock-control none --import-source on, {args.command}".strip(", "),
}
json.dump(doc, open(args.out, "w"), indent=1)
print(json.dumps(doc, indent=1))
