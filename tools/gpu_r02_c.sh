#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_pipeline_pool.py tests/test_pifpaf.py "tests/test_engine_gpu.py::test_resnet50_pifpaf_fields_and_decode" -x -q > gpurun_out/r02c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_tests.log)
tail -12 gpurun_out/r02c_tests.log
for r in default 0 8 24; do
  if [ "$r" = default ]; then unset HPB_PIFPAF_RESERVE_SMS; else export HPB_PIFPAF_RESERVE_SMS=$r; fi
  timeout 300 python bench.py --workload cfg5 --steps 20 --no-cpu-baseline > gpurun_out/r02c_bench_cfg5_res_$r.json 2> gpurun_out/r02c_bench_cfg5_res_$r.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02c_bench_cfg5_res_$r.json"))
    print("cfg5 reserve=$r value %.1f e2e %.1f sync %.1f ms/step %.3f conv_ms %.3f parse_ms %.3f"%(d["value"],d["e2e"]["value"],d["e2e"]["synchronous_call"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"],d["breakdown"]["parse_ms_per_step"]))
except Exception as ex: print("cfg5 reserve=$r failed",ex); print(open("gpurun_out/r02c_bench_cfg5_res_$r.err").read()[-1500:])
PY
done
unset HPB_PIFPAF_RESERVE_SMS
timeout 300 python bench.py --no-extra --no-tf32-line --no-cpu-baseline --steps 30 > gpurun_out/r02c_bench_cfg3.json 2> gpurun_out/r02c_bench_cfg3.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02c_bench_cfg3.json"))
print("cfg3 value %.1f e2e %.1f ms/step %.3f conv_ms %.3f frac %.3f parse_ms %.4f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"],d["roofline"]["frac"],d["breakdown"]["parse_ms_per_step"]))
print(d["breakdown"].get("parse_alone_by_batch"))
PY
