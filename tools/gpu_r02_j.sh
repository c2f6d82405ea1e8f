#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_backbone_fullsize.py -x -q > gpurun_out/r02j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02j_tests.log); tail -12 gpurun_out/r02j_tests.log
run() { name=$1; k=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $k --steps 30 --no-cpu-baseline > gpurun_out/r02j_bench_${k}_$name.json 2> gpurun_out/r02j_bench_${k}_$name.err
  cp gpurun_out/bench_layers_${k}_f16_n1.json gpurun_out/r02j_layers_${k}_$name.json; }
run fused cfg2 X=1
run plain cfg2 HPB_NO_DW1_FUSE=1 HPB_NO_DW_DUAL=1
run fused2 cfg2 X=1
python - <<PY
import json
for v in ("fused","plain","fused2"):
    try:
        d=json.load(open("gpurun_out/r02j_bench_cfg2_%s.json"%v)); b=d["breakdown"]
        print(v,"cfg2 value %.1f e2e %.1f ms/step %.3f conv_ms %.3f backbone_ms %.3f launches %d"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"],b["backbone_ms_per_step"],d["gpu_launches"]))
    except Exception as ex: print(v,"failed",ex, open("gpurun_out/r02j_bench_cfg2_%s.err"%v).read()[-800:])
PY
