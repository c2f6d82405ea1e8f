#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_backbone_fullsize.py -x -q > gpurun_out/r02g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02g_tests.log); tail -4 gpurun_out/r02g_tests.log
for k in cfg3 cfg4; do
  extra=""; [ $k = cfg3 ] && extra="--no-extra --no-tf32-line"
  timeout 300 python bench.py --workload $k $extra --steps 30 --no-cpu-baseline > gpurun_out/r02g_bench_$k.json 2> gpurun_out/r02g_bench_$k.err
  cp gpurun_out/bench_layers_${k}_f16_n1.json gpurun_out/r02g_layers_$k.json
done
python - <<PY
import json
for k in ("cfg3","cfg4"):
    d=json.load(open("gpurun_out/r02g_bench_%s.json"%k))
    print(k,"value %.1f e2e %.1f ms/step %.3f conv_ms %.3f frac %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms_per_step"],d["roofline"]["frac"]))
L={l["name"]:l["ms"] for l in json.load(open("gpurun_out/r02g_layers_cfg3.json"))["layers"]}
print({k:round(L[k],4) for k in ("conv1_1","conv1_2","conv2_1","init_4","init_out","ref1_6","ref1_out","ref5_out","conv3_2","ref1_1","cpm_1","conv4_2")})
PY
# cfg2: where does the depthwise time go
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02g_cfg2_launches.csv python tools/profile_cfg.py --steps 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv3_col -s 40 -c 2 -f -o gpurun_out/r02g_dw python tools/profile_cfg.py --steps 1 > gpurun_out/r02g_ncu_dw.log 2>&1
ls -la gpurun_out/r02g_dw.ncu-rep
