#!/bin/bash
# dwconv3_tma_kernel parameter sweep on cfg2 (tile order / tile height / threads), PDL on by rule; then the full GPU suite
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --workload cfg2 --steps 30 --no-cpu-baseline > gpurun_out/r02r_bench_cfg2_$name.json 2> gpurun_out/r02r_bench_cfg2_$name.err
  cp gpurun_out/bench_layers_cfg2_f16_n1.json gpurun_out/r02r_layers_cfg2_$name.json; }
run A X=1
run B HPB_DWT_ORDER=1
run C HPB_DWT_ORDER=1 HPB_DWT_HB=8
run D HPB_DWT_ORDER=1 HPB_DWT_HB=8 HPB_DWT_THREADS=384
run E HPB_DWT_ORDER=1 HPB_DWT_THREADS=384
run F HPB_DWT_THREADS=384
run G HPB_DWT_ORDER=1 HPB_DWT_HB=4
run H HPB_DWT_ORDER=1 HPB_DWT_HB=8 HPB_DWT_THREADS=384 HPB_PDL=0
run I HPB_DWT_ORDER=1 HPB_DWT_HB=8 HPB_DWT_THREADS=480
run A2 X=1
python - <<PY
import json
for v in ("A","B","C","D","E","F","G","H","I","A2"):
    try:
        d=json.load(open("gpurun_out/r02r_bench_cfg2_%s.json"%v))
        L=json.load(open("gpurun_out/r02r_layers_cfg2_%s.json"%v))["layers"]
        dw=[l for l in L if "_dw" in l["name"]]
        print("cfg2",v,"value %.1f e2e %.1f ms/step %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"]), "dw total %.4f"%sum(l["ms"] for l in dw), {l["name"]:round(l["ms"],4) for l in dw if l["name"] in ("convblock_3_dw","convblock_5_dw","convblock_7_dw","init_1_dw0","init_2_dw")})
    except Exception as ex: print(v,"failed",ex)
PY
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02r_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/r02r_gputests.log); tail -n 4 gpurun_out/r02r_gputests.log
