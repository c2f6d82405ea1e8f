#!/bin/bash
# halo MMA issuer (resident / streamed instantiations) old vs new, and the hybrid N-half tile list (HPB_SPLIT=1): tests, then same-box A/B
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_backbone_fullsize.py -x -q > gpurun_out/r02o_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02o_tests.log); tail -3 gpurun_out/r02o_tests.log
(HPB_SPLIT=1 timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_backbone_fullsize.py tests/test_pipeline_pool.py -x -q > gpurun_out/r02o_tests_split.log 2>&1; echo "rc=$?" >> gpurun_out/r02o_tests_split.log); tail -3 gpurun_out/r02o_tests_split.log
run() { name=$1; k=$2; shift 2
  extra=""; [ $k = cfg3 ] && extra="--no-extra --no-tf32-line"
  env "$@" timeout 300 python bench.py --workload $k $extra --steps 30 --no-cpu-baseline > gpurun_out/r02o_bench_${k}_$name.json 2> gpurun_out/r02o_bench_${k}_$name.err
  cp gpurun_out/bench_layers_${k}_f16_n1.json gpurun_out/r02o_layers_${k}_$name.json; }
for rep in 1 2; do
  cp hyperpose_b200/lib_new_halo.so hyperpose_b200/libhyperpose_b200.so
  for k in cfg3 cfg4 cfg5; do run new$rep $k X=1; run split$rep $k HPB_SPLIT=1; done
  cp hyperpose_b200/lib_old_halo.so hyperpose_b200/libhyperpose_b200.so
  for k in cfg3 cfg4 cfg5; do run old$rep $k X=1; done
done
cp hyperpose_b200/lib_new_halo.so hyperpose_b200/libhyperpose_b200.so
python - <<PY
import json
for k in ("cfg3","cfg4","cfg5"):
    for v in ("new1","split1","old1","new2","split2","old2"):
        try:
            d=json.load(open("gpurun_out/r02o_bench_%s_%s.json"%(k,v)))
            print(k,v,"value %.1f e2e %.1f ms/step %.3f conv_ms %.3f frac %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"],d["roofline"]["frac"]))
        except Exception as ex: print(k,v,"failed",ex)
for v in ("new1","split1","old1"):
    L={l["name"]:l["ms"] for l in json.load(open("gpurun_out/r02o_layers_cfg3_%s.json"%v))["layers"]}
    print(v,{k:round(L[k],4) for k in ("conv1_2","conv2_1","conv2_2","conv3_2","conv4_2","cpm_1","init_1","ref1_1","ref1_2","ref3_3","init_2") if k in L})
PY
