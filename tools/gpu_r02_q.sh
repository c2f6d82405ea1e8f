#!/bin/bash
# dwconv3_tma_kernel v2 (576 threads, filters reloaded on change, 6-row tiles x 3 stages): tests, cfg2 A/B incl. HPB_PDL, ncu launch list
mkdir -p gpurun_out
(timeout 500 python -m pytest tests/test_engine_gpu.py -x -q -k "tma_tiled or n_half or mobilenet or fused_depthwise" > gpurun_out/r02q_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02q_tests.log); tail -n 4 gpurun_out/r02q_tests.log
run() { name=$1; k=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $k --steps 30 --no-cpu-baseline > gpurun_out/r02q_bench_${k}_$name.json 2> gpurun_out/r02q_bench_${k}_$name.err
  cp gpurun_out/bench_layers_${k}_f16_n1.json gpurun_out/r02q_layers_${k}_$name.json; }
for rep in 1 2; do run tma$rep cfg2 X=1; run plain$rep cfg2 HPB_NO_DW_TMA=1; run pdl$rep cfg2 HPB_PDL=1; done
python - <<PY
import json
for v in ("tma1","plain1","pdl1","tma2","plain2","pdl2"):
    try:
        d=json.load(open("gpurun_out/r02q_bench_cfg2_%s.json"%v))
        print("cfg2",v,"value %.1f e2e %.1f ms/step %.3f conv_ms %.3f frac %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"],d["roofline"]["frac"]))
        L=json.load(open("gpurun_out/r02q_layers_cfg2_%s.json"%v))["layers"]
        dw=[l for l in L if "_dw" in l["name"]]
        print("   dw total %.4f ms; "%sum(l["ms"] for l in dw), {l["name"]:round(l["ms"],4) for l in dw if l["name"] in ("convblock_1_dw","convblock_3_dw","convblock_5_dw","convblock_7_dw","convblock_11_dw","init_1_dw0","init_2_dw","ref1_1_dw0","ref1_2_dw","ref5_3_dw")})
    except Exception as ex: print(v,"failed",ex)
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02q_launches_cfg2.csv python tools/profile_cfg.py --steps 2 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwconv3_tma -s 20 -c 3 -f -o gpurun_out/r02q_dwtma python tools/profile_cfg.py --steps 1 > gpurun_out/r02q_ncu.log 2>&1
ls -la gpurun_out/r02q_dwtma.ncu-rep
