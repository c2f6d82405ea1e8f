#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; k=$2; shift 2
  extra=""; [ $k = cfg3 ] && extra="--no-extra --no-tf32-line"
  timeout 300 python bench.py --workload $k $extra --steps 30 --no-cpu-baseline > gpurun_out/r02m_bench_${k}_$name.json 2> gpurun_out/r02m_bench_${k}_$name.err
  cp gpurun_out/bench_layers_${k}_f16_n1.json gpurun_out/r02m_layers_${k}_$name.json; }
for rep in 1 2; do
  cp hyperpose_b200/lib_new_issue.so hyperpose_b200/libhyperpose_b200.so; run new$rep cfg3; run new$rep cfg4
  cp hyperpose_b200/lib_old_issue.so hyperpose_b200/libhyperpose_b200.so; run old$rep cfg3; run old$rep cfg4
done
cp hyperpose_b200/lib_new_issue.so hyperpose_b200/libhyperpose_b200.so
python - <<PY
import json
for v in ("new1","old1","new2","old2"):
    for k in ("cfg3","cfg4"):
        d=json.load(open("gpurun_out/r02m_bench_%s_%s.json"%(k,v)))
        print(v,k,"value %.1f e2e %.1f ms/step %.3f conv_ms %.3f frac %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"],d["roofline"]["frac"]))
    L={l["name"]:l["ms"] for l in json.load(open("gpurun_out/r02m_layers_cfg3_%s.json"%v))["layers"]}
    print("   ",{k:round(L[k],4) for k in ("conv3_2","conv4_2","cpm_1","ref1_1","ref1_2","ref3_3","init_2","ref1_6","init_4")})
PY
(timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_backbone_fullsize.py -x -q > gpurun_out/r02m_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02m_tests.log); tail -3 gpurun_out/r02m_tests.log
