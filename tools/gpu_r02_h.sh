#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_backbone_fullsize.py tests/test_pipeline_pool.py tests/test_handoff.py tests/test_engine_tf32.py -x -q > gpurun_out/r02h_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02h_tests.log); tail -4 gpurun_out/r02h_tests.log
for v in pdl nopdl pdl2; do
  if [ "$v" = nopdl ]; then export HPB_NO_PDL=1; else unset HPB_NO_PDL; fi
  for k in cfg3 cfg4; do
    extra=""; [ $k = cfg3 ] && extra="--no-extra --no-tf32-line"
    timeout 300 python bench.py --workload $k $extra --steps 30 --no-cpu-baseline > gpurun_out/r02h_bench_${k}_$v.json 2> gpurun_out/r02h_bench_${k}_$v.err
  done
done
unset HPB_NO_PDL
python - <<PY
import json
for v in ("pdl","nopdl","pdl2"):
    for k in ("cfg3","cfg4"):
        try:
            d=json.load(open("gpurun_out/r02h_bench_%s_%s.json"%(k,v)))
            print(v,k,"value %.1f e2e %.1f ms/step %.3f conv_ms %.3f frac %.3f graphs %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"],d["roofline"]["frac"],d["e2e"]["graphs"]))
        except Exception as ex: print(v,k,"failed",ex, open("gpurun_out/r02h_bench_%s_%s.err"%(k,v)).read()[-800:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv3_col -s 12 -c 2 -f -o gpurun_out/r02h_dw python tools/profile_cfg.py --steps 1 > gpurun_out/r02h_ncu_dw.log 2>&1
ls -la gpurun_out/r02h_dw.ncu-rep
