#!/bin/bash
# lean final run with the one-barrier epilogue as the default: full GPU suite + the default bench line
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final3_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final3_gputests.log); tail -4 gpurun_out/r02_final3_gputests.log
timeout 500 python bench.py > gpurun_out/r02_final3_bench_default.json 2> gpurun_out/r02_final3_bench_default.err; echo "bench rc=$?"
for k in cfg2 cfg3 cfg4 cfg5; do cp gpurun_out/bench_layers_${k}_f16_n1.json gpurun_out/r02_final3_layers_$k.json; done
cp gpurun_out/bench_layers_cfg3_tf32_n1.json gpurun_out/r02_final3_layers_cfg3_tf32.json
python - <<PY
import json
d=json.load(open("gpurun_out/r02_final3_bench_default.json"))
def show(x,name):
    r=x["roofline"]; b=x.get("breakdown",{})
    print(name,"value %.1f e2e %.1f ms/step %.3f conv_ms %.3f frac %.3f parse_ms %s"%(x["value"],x["e2e"]["value"],x["ms_per_step"],r["kernel_ms_per_step"],r["frac"],b.get("parse_ms_per_step")))
show(d,"cfg3"); print(d["clocks"])
for e in d["extra_configs"]:
    if "error" in e: print(e)
    else: show(e,e["metric"][-60:])
show(d["tf32"],"tf32") if "error" not in d["tf32"] else print(d["tf32"])
PY
