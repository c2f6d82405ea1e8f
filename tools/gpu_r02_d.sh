#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_backbone_fullsize.py tests/test_engine_tf32.py -x -q -k "resnet50 or pifpaf" > gpurun_out/r02d_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02d_tests.log)
tail -12 gpurun_out/r02d_tests.log
for v in v2 v1; do
  if [ "$v" = v1 ]; then export HPB_STEM7_V1=1; else unset HPB_STEM7_V1; fi
  timeout 300 python bench.py --workload cfg4 --steps 20 --no-cpu-baseline > gpurun_out/r02d_bench_cfg4_$v.json 2> gpurun_out/r02d_bench_cfg4_$v.err
  cp gpurun_out/bench_layers_cfg4_f16_n1.json gpurun_out/r02d_layers_cfg4_$v.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02d_bench_cfg4_$v.json")); L=json.load(open("gpurun_out/r02d_layers_cfg4_$v.json"))["layers"]
    print("cfg4 stem7 $v value %.1f e2e %.1f ms/step %.3f conv_ms %.3f frac %.3f stem %.4f ms"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"],d["roofline"]["frac"],[l["ms"] for l in L if l["name"]=="conv1+bn1"][0]))
except Exception as ex: print("cfg4 $v failed",ex); print(open("gpurun_out/r02d_bench_cfg4_$v.err").read()[-1500:])
PY
done
unset HPB_STEM7_V1
timeout 300 python bench.py --workload cfg5 --steps 20 --no-cpu-baseline > gpurun_out/r02d_bench_cfg5.json 2> gpurun_out/r02d_bench_cfg5.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02d_bench_cfg5.json")); print("cfg5 value %.1f e2e %.1f ms/step %.3f conv_ms %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"]))
PY
