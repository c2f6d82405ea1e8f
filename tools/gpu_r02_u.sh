#!/bin/bash
# one-barrier form of the TMA-store epilogue (HPB_EPI_1BAR=1): tests under it, then same-box A/B on cfg4 / cfg5 / cfg3 / cfg2
mkdir -p gpurun_out
(HPB_EPI_1BAR=1 timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_backbone_fullsize.py -x -q > gpurun_out/r02u_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02u_tests.log); tail -n 4 gpurun_out/r02u_tests.log
run() { name=$1; k=$2; shift 2
  extra=""; [ $k = cfg3 ] && extra="--no-extra --no-tf32-line"
  env "$@" timeout 300 python bench.py --workload $k $extra --steps 30 --no-cpu-baseline > gpurun_out/r02u_bench_${k}_$name.json 2> gpurun_out/r02u_bench_${k}_$name.err
  cp gpurun_out/bench_layers_${k}_f16_n1.json gpurun_out/r02u_layers_${k}_$name.json; }
for k in cfg4 cfg5 cfg3 cfg2; do run one $k HPB_EPI_1BAR=1; run two $k HPB_EPI_1BAR=0; done
run one2 cfg4 HPB_EPI_1BAR=1; run two2 cfg4 HPB_EPI_1BAR=0
python - <<PY
import json
for k,v in (("cfg4","one"),("cfg4","two"),("cfg4","one2"),("cfg4","two2"),("cfg5","one"),("cfg5","two"),("cfg3","one"),("cfg3","two"),("cfg2","one"),("cfg2","two")):
    try:
        d=json.load(open("gpurun_out/r02u_bench_%s_%s.json"%(k,v)))
        L={l["name"]:l["ms"] for l in json.load(open("gpurun_out/r02u_layers_%s_%s.json"%(k,v)))["layers"]}
        print(k,v,"value %.1f e2e %.1f ms/step %.3f conv_ms %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"]), {n:round(L[n],4) for n in ("block_1_2_conv3","block_2_2_conv3","block_3_2_conv3","block_4_2_conv3","ref1_6","ref1_out","init_4") if n in L})
    except Exception as ex: print(k,v,"failed",ex)
PY
