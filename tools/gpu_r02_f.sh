#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_backbone_fullsize.py -x -q > gpurun_out/r02f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02f_tests.log); tail -4 gpurun_out/r02f_tests.log
for v in epi8 epi4; do
  if [ "$v" = epi4 ]; then export HPB_EPI4=1; else unset HPB_EPI4; fi
  for k in cfg4 cfg3; do
    extra=""; [ $k = cfg3 ] && extra="--no-extra --no-tf32-line"
    timeout 300 python bench.py --workload $k $extra --steps 30 --no-cpu-baseline > gpurun_out/r02f_bench_${k}_$v.json 2> gpurun_out/r02f_bench_${k}_$v.err
    cp gpurun_out/bench_layers_${k}_f16_n1.json gpurun_out/r02f_layers_${k}_$v.json
  done
done
unset HPB_EPI4
python - <<PY
import json
for v in ("epi8","epi4"):
    for k in ("cfg4","cfg3"):
        try:
            d=json.load(open("gpurun_out/r02f_bench_%s_%s.json"%(k,v)))
            print(v,k,"value %.1f e2e %.1f ms/step %.3f conv_ms %.3f frac %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"],d["roofline"]["frac"]))
        except Exception as ex: print(v,k,"failed",ex)
    L={l["name"]:l["ms"] for l in json.load(open("gpurun_out/r02f_layers_cfg4_%s.json"%v))["layers"]}
    print(v,{k:round(L[k],4) for k in ("block_4_1_conv3","block_3_1_conv3","block_1_1_conv3","block_4_1_conv2","block_4_1_ds")})
    L={l["name"]:l["ms"] for l in json.load(open("gpurun_out/r02f_layers_cfg3_%s.json"%v))["layers"]}
    print(v,{k:round(L[k],4) for k in ("init_4","init_out","ref1_6","ref1_out","ref5_out","conv3_2","ref1_1","cpm_1","conv4_2")})
PY
