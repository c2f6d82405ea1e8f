#!/bin/bash
# 8 GPUs of one box: frame-sharded cfg4 (BASELINE config 4: batch 32 per GPU x 8)
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --workload cfg4 --no-cpu-baseline > gpurun_out/r02_bench_cfg4_n8.json 2> gpurun_out/r02_bench_cfg4_n8.err; echo "cfg4 n8 rc=$?"
python - <<PY
import json
for f in ("r02_bench_cfg4_n8",):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f,"n_gpus",d["n_gpus"],"value %.1f e2e %.1f ms/step %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"]))
    except Exception as ex: print(f,"failed",ex, open("gpurun_out/%s.err"%f).read()[-600:])
PY

