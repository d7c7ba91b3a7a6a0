import json, sys
d=json.load(open(sys.argv[1]))
print("HEAD", d["metric"], round(d["value"]), d["ms_per_step"], d.get("step_time_stats"), "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "e2e", round(d["e2e"]["value"]), "launches", d.get("gpu_launches"))
for k,v in d["kernel_classes"].items(): print("  ",k, {a:round(b,3) for a,b in v.items()})
print("cpu", d.get("cpu_baseline"), "clocks", d.get("clocks"))
for k,v in d.get("extra_configs",{}).items():
    print(k, {a:v.get(a) for a in ("value","unit","ms_per_step","error")}, "frac", (v.get("roofline") or {}).get("frac"), "cpu", (v.get("cpu_baseline") or {}).get("value"), "e2e", (v.get("e2e") or {}).get("value"))
    for kk,vv in (v.get("kernel_classes") or {}).items(): print("     ",kk, {a:round(b,3) for a,b in vv.items()})
    if "kernels" in v:
        for r in v["kernels"]: print("    ", {a:(round(b,3) if isinstance(b,float) else b) for a,b in r.items()})
