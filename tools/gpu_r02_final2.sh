#!/bin/bash
# final run of the round: full GPU suite, smoke, the default bench line, the reference arm, launch lists of cfg2 / cfg3 with the final code
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final2_gputests.log); tail -5 gpurun_out/r02_final2_gputests.log
(timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_final2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_final2_smoke.log); tail -2 gpurun_out/r02_final2_smoke.log
timeout 600 python bench.py > gpurun_out/r02_final2_bench_default.json 2> gpurun_out/r02_final2_bench_default.err; echo "bench rc=$?"
for k in cfg2 cfg3 cfg4 cfg5; do cp gpurun_out/bench_layers_${k}_f16_n1.json gpurun_out/r02_final2_layers_$k.json; done
cp gpurun_out/bench_layers_cfg3_tf32_n1.json gpurun_out/r02_final2_layers_cfg3_tf32.json
timeout 400 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_final2_bench_reference.json 2> gpurun_out/r02_final2_bench_reference.err; echo "reference arm rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r02_final2_bench_default.json"))
def show(x,name):
    r=x["roofline"]; b=x.get("breakdown",{})
    print(name,"value %.1f e2e %.1f ms/step %.3f conv_ms %.3f frac %.3f parse_ms %s"%(x["value"],x["e2e"]["value"],x["ms_per_step"],r["kernel_ms_per_step"],r["frac"],b.get("parse_ms_per_step")))
show(d,"cfg3"); print(d["clocks"]); print(d.get("cpu_baseline",{}).get("value"), d.get("parse_only"))
for e in d["extra_configs"]:
    if "error" in e: print(e)
    else: show(e,e["metric"][-60:])
show(d["tf32"],"tf32") if "error" not in d["tf32"] else print(d["tf32"])
try: print(open("gpurun_out/r02_final2_bench_reference.json").read()[:600])
except Exception as ex: print(ex)
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_final2_launches_step.csv python tools/profile_step.py --steps 2 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_final2_launches_cfg2.csv python tools/profile_cfg.py --steps 2 > /dev/null 2>&1
ls -la gpurun_out/r02_final2_launches_*.csv
