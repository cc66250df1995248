"""Print the headline fields of a bench.py JSON line:  python tools/bench_show.py FILE   (value, e2e, verified, phases, launches, ms/step, extras)"""
import json,sys
d=json.loads([x for x in open(sys.argv[1]) if x.startswith("{")][-1])
print(d["value"], d["e2e"]["value"], d.get("verified"), d.get("phases_ms_per_step"), d["gpu_launches"], d["ms_per_step"])
for k,v in d.get("extras",{}).items(): print(k, {a:b for a,b in v.items() if a in ("frame_ms","step_ms","analyzer_ms","job_ms","kernel_launches","checksum_equals_config4_strong","Mcell_updates_per_s")})
