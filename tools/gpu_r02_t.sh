#!/bin/bash
# residual epilogue: leader decodes the residual tile once per tile, residual tiles read with ld.shared -- tests + cfg4 / cfg5 / cfg2 bench
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_backbone_fullsize.py -x -q > gpurun_out/r02t_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02t_tests.log); tail -n 4 gpurun_out/r02t_tests.log
run() { name=$1; k=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $k --steps 30 --no-cpu-baseline > gpurun_out/r02t_bench_${k}_$name.json 2> gpurun_out/r02t_bench_${k}_$name.err
  cp gpurun_out/bench_layers_${k}_f16_n1.json gpurun_out/r02t_layers_${k}_$name.json; }
run a cfg4 X=1; run a cfg5 X=1; run a cfg2 X=1; run b cfg4 X=1
python - <<PY
import json
for k,v in (("cfg4","a"),("cfg5","a"),("cfg2","a"),("cfg4","b")):
    try:
        d=json.load(open("gpurun_out/r02t_bench_%s_%s.json"%(k,v)))
        L={l["name"]:l["ms"] for l in json.load(open("gpurun_out/r02t_layers_%s_%s.json"%(k,v)))["layers"]}
        print(k,v,"value %.1f e2e %.1f ms/step %.3f conv_ms %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"]), {n:round(L[n],4) for n in ("block_1_2_conv3","block_2_2_conv3","block_3_2_conv3","block_4_2_conv3") if n in L})
    except Exception as ex: print(k,v,"failed",ex)
PY
