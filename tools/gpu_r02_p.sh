#!/bin/bash
# dwconv3_tma_kernel: bit-identity tests, then cfg2 same-box A/B against the per-lane-load kernels
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_engine_gpu.py -x -q -k "tma_tiled or n_half or mobilenet or fused_depthwise" > gpurun_out/r02p_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02p_tests.log); tail -n 15 gpurun_out/r02p_tests.log
(timeout 400 python -m pytest tests/test_backbone_fullsize.py -x -q > gpurun_out/r02p_tests_full.log 2>&1; echo "rc=$?" >> gpurun_out/r02p_tests_full.log); tail -n 3 gpurun_out/r02p_tests_full.log
run() { name=$1; k=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $k --steps 30 --no-cpu-baseline > gpurun_out/r02p_bench_${k}_$name.json 2> gpurun_out/r02p_bench_${k}_$name.err
  cp gpurun_out/bench_layers_${k}_f16_n1.json gpurun_out/r02p_layers_${k}_$name.json; }
for rep in 1 2; do run tma$rep cfg2 X=1; run plain$rep cfg2 HPB_NO_DW_TMA=1; done
python - <<PY
import json
for v in ("tma1","plain1","tma2","plain2"):
    try:
        d=json.load(open("gpurun_out/r02p_bench_cfg2_%s.json"%v))
        print("cfg2",v,"value %.1f e2e %.1f ms/step %.3f conv_ms %.3f frac %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"],d["roofline"]["frac"]))
        L=json.load(open("gpurun_out/r02p_layers_cfg2_%s.json"%v))["layers"]
        dw=[l for l in L if "_dw" in l["name"]]
        print("   dw total %.4f ms; "%sum(l["ms"] for l in dw), {l["name"]:round(l["ms"],4) for l in dw if l["name"] in ("convblock_1_dw","convblock_3_dw","convblock_5_dw","convblock_7_dw","convblock_11_dw","init_1_dw0","init_2_dw","ref1_1_dw0","ref1_2_dw","ref5_3_dw")})
    except Exception as ex: print(v,"failed",ex)
PY
