#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; k=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $k --no-extra --no-tf32-line --steps 30 --no-cpu-baseline > gpurun_out/r02l_bench_${k}_$name.json 2> gpurun_out/r02l_bench_${k}_$name.err
  cp gpurun_out/bench_layers_${k}_f16_n1.json gpurun_out/r02l_layers_${k}_$name.json; }
run base cfg3 X=1
run halopool cfg3 HPB_HALO_POOL=1
run st3 cfg3 HPB_SWAP_STAGES=3
run st2 cfg3 HPB_SWAP_STAGES=2
run base2 cfg3 X=1
python - <<PY
import json
for v in ("base","halopool","st3","st2","base2"):
    try:
        d=json.load(open("gpurun_out/r02l_bench_cfg3_%s.json"%v)); b=d["breakdown"]
        print(v,"cfg3 value %.1f e2e %.1f ms/step %.3f conv_ms %.3f backbone_ms %.3f frac %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"],b["backbone_ms_per_step"],d["roofline"]["frac"]))
        L={l["name"]:l["ms"] for l in json.load(open("gpurun_out/r02l_layers_cfg3_%s.json"%v))["layers"]}
        print("   ",{k:round(L[k],4) for k in ("conv1_2","conv2_1","conv2_2","maxpool_2","conv3_2","ref1_1","ref1_2","ref3_3","init_2","cpm_2")})
    except Exception as ex: print(v,"failed",ex)
PY
HPB_HALO_POOL=1 timeout 300 python -m pytest tests/test_backbone_fullsize.py -x -q -k vgg19 2>&1 | tail -3
