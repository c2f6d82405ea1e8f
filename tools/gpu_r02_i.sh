#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_backbone_fullsize.py tests/test_engine_tf32.py tests/test_pipeline_pool.py -x -q > gpurun_out/r02i_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02i_tests.log); tail -6 gpurun_out/r02i_tests.log
run() { # name workload envs...
  name=$1; k=$2; shift 2
  extra=""; [ $k = cfg3 ] && extra="--no-extra --no-tf32-line"
  env "$@" timeout 300 python bench.py --workload $k $extra --steps 30 --no-cpu-baseline > gpurun_out/r02i_bench_${k}_$name.json 2> gpurun_out/r02i_bench_${k}_$name.err
  cp gpurun_out/bench_layers_${k}_f16_n1.json gpurun_out/r02i_layers_${k}_$name.json
}
run dual cfg2 X=1
run nodual cfg2 HPB_NO_DW_DUAL=1
run fuse cfg3 X=1
run nofuse cfg3 HPB_NO_POOL_FUSE=1
run fuse2 cfg3 X=1
python - <<PY
import json
for k,v in (("cfg2","dual"),("cfg2","nodual"),("cfg3","fuse"),("cfg3","nofuse"),("cfg3","fuse2")):
    try:
        d=json.load(open("gpurun_out/r02i_bench_%s_%s.json"%(k,v)))
        b=d["breakdown"]
        print(v,k,"value %.1f e2e %.1f ms/step %.3f conv_ms %.3f backbone_ms %.3f frac %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"],b["backbone_ms_per_step"],d["roofline"]["frac"]))
    except Exception as ex: print(v,k,"failed",ex, open("gpurun_out/r02i_bench_%s_%s.err"%(k,v)).read()[-800:])
for v in ("fuse","nofuse"):
    L={l["name"]:l["ms"] for l in json.load(open("gpurun_out/r02i_layers_cfg3_%s.json"%v))["layers"]}
    print(v,{k:round(L[k],4) for k in ("conv1_1","conv1_2","maxpool_1","conv2_1","conv2_2","maxpool_2")})
PY
