#!/bin/bash
# residual-tile ring of depth 4 for the short-k residual layers (ResNet conv3): tests, cfg4 / cfg5 same-box A/B, ncu of the residual kernel
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_backbone_fullsize.py -x -q > gpurun_out/r02s_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02s_tests.log); tail -n 4 gpurun_out/r02s_tests.log
run() { name=$1; k=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $k --steps 30 --no-cpu-baseline > gpurun_out/r02s_bench_${k}_$name.json 2> gpurun_out/r02s_bench_${k}_$name.err
  cp gpurun_out/bench_layers_${k}_f16_n1.json gpurun_out/r02s_layers_${k}_$name.json; }
for rep in 1 2; do
  for k in cfg4 cfg5; do run res4_$rep $k X=1; run res2_$rep $k HPB_RES_STAGES=2; run res4k8_$rep $k HPB_RES4_KMAX=8; done
done
python - <<PY
import json
for k in ("cfg4","cfg5"):
    for v in ("res4_1","res2_1","res4k8_1","res4_2","res2_2","res4k8_2"):
        try:
            d=json.load(open("gpurun_out/r02s_bench_%s_%s.json"%(k,v)))
            L={l["name"]:l["ms"] for l in json.load(open("gpurun_out/r02s_layers_%s_%s.json"%(k,v)))["layers"]}
            print(k,v,"value %.1f e2e %.1f ms/step %.3f conv_ms %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"]), {n:round(L[n],4) for n in ("block_1_2_conv3","block_2_2_conv3","block_3_2_conv3","block_4_2_conv3") if n in L})
        except Exception as ex: print(k,v,"failed",ex)
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tcgen05_kernel -s 8 -c 14 -f -o gpurun_out/r02s_res python tools/profile_cfg.py --graph resnet50_lw_openpose --batch 32 --steps 1 > gpurun_out/r02s_ncu.log 2>&1
ls -la gpurun_out/r02s_res.ncu-rep
