#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload cfg4 --steps 20 --no-cpu-baseline > gpurun_out/r02e_bench_cfg4.json 2> gpurun_out/r02e_bench_cfg4.err
cp gpurun_out/bench_layers_cfg4_f16_n1.json gpurun_out/r02e_layers_cfg4.json
timeout 300 python bench.py --no-extra --no-tf32-line --no-cpu-baseline --steps 30 > gpurun_out/r02e_bench_cfg3.json 2> gpurun_out/r02e_bench_cfg3.err
cp gpurun_out/bench_layers_cfg3_f16_n1.json gpurun_out/r02e_layers_cfg3.json
python - <<PY
import json
for k in ("cfg4","cfg3"):
    d=json.load(open("gpurun_out/r02e_bench_%s.json"%k))
    print(k,"value %.1f e2e %.1f ms/step %.3f conv_ms %.3f frac %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"],d["roofline"]["frac"]))
L={l["name"]:l["ms"] for l in json.load(open("gpurun_out/r02e_layers_cfg4.json"))["layers"]}
print({k:round(L[k],4) for k in ("block_4_1_conv3","block_3_1_conv3","block_1_1_conv3","conv1+bn1","block_4_1_conv2")})
L={l["name"]:l["ms"] for l in json.load(open("gpurun_out/r02e_layers_cfg3.json"))["layers"]}
print({k:round(L[k],4) for k in ("conv1_2","conv2_1","init_4","init_out","ref1_6","ref1_out","ref5_out","conv3_2","ref1_1","ref1_2")})
PY
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_gputests.log); tail -5 gpurun_out/r02e_gputests.log
