"""Folds an ncu CSV launch list (one row per kernel launch and metric: gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum; `--clock-control none`, captured over tools/profile_step.py) into
  * profiles/conv_traffic.json  -- DRAM bytes of the conv launches of ONE step (bench.py's roofline.traffic)
  * a per-kernel-name table (launches, total time, share of the step) printed as text.
usage: python tools/summarize_launches.py <ncu.csv> <steps in the capture> [--write]"""
import csv
import json
import os
import sys
from collections import OrderedDict

path, steps = sys.argv[1], int(sys.argv[2])
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)
launch = OrderedDict()
for r in rows:
    d = launch.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0].replace("void ", "").replace("hpb::", ""), "grid": r["Grid Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    else:
        d[r["Metric Name"]] = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
ls = list(launch.values())
# the last `1/steps` of the launches = one warm step (the first step pays cold caches and lazy initialisation)
per = len(ls) // steps
step = ls[-per:]
tot = sum(l.get("us", 0.0) for l in step)
by = OrderedDict()
for l in step:
    b = by.setdefault(l["name"], {"launches": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
    b["launches"] += 1; b["us"] += l.get("us", 0.0); b["rd"] += l.get("dram__bytes_read.sum", 0.0); b["wr"] += l.get("dram__bytes_write.sum", 0.0)
print(f"{len(ls)} launches captured, {per} per step; device time of the last step {tot / 1e3:.3f} ms (ncu-serialised, cold caches)")
print(f"{'kernel':48s} {'n':>4s} {'ms':>8s} {'share':>7s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s}")
for k, b in sorted(by.items(), key=lambda kv: -kv[1]["us"]):
    print(f"{k[:48]:48s} {b['launches']:4d} {b['us'] / 1e3:8.3f} {b['us'] / tot:7.1%} {b['rd'] / 1e6:11.1f} {b['wr'] / 1e6:11.1f}")
conv = [b for k, b in by.items() if k.startswith("conv_")]
out = {"dram_bytes_per_step": sum(b["rd"] + b["wr"] for b in conv), "dram_read_bytes": sum(b["rd"] for b in conv),
       "dram_write_bytes": sum(b["wr"] for b in conv), "conv_launches": sum(b["launches"] for b in conv),
       "conv_share_of_step_device_time": sum(b["us"] for b in conv) / tot,
       "source": "ncu capture " + os.path.basename(path) + " committed under profiles/ (not measured in the bench run itself)",
       "note": "sum over the conv launches (conv_tcgen05 / swap / halo / stem kernels) of one warm cfg3 step (batch 16): ncu --metrics "
               "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over tools/profile_step.py (" + os.path.basename(path) + ")"}
print(json.dumps(out, indent=1))
if "--write" in sys.argv:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    json.dump(out, open(os.path.join(root, "profiles", "conv_traffic.json"), "w"), indent=1)
