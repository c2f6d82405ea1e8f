#!/bin/bash
# 2 GPUs of one box: frame-sharded cfg4 (BASELINE config 4: batch 32 x 8) and cfg3, and the in-process pool at world size 2
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --workload cfg4 --no-cpu-baseline > gpurun_out/r02_bench_cfg4_n2.json 2> gpurun_out/r02_bench_cfg4_n2.err; echo "cfg4 n8 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_cfg3_n2.json 2> gpurun_out/r02_bench_cfg3_n2.err; echo "cfg3 n8 rc=$?"
timeout 300 python bench.py --workload cfg4 --steps 20 --no-cpu-baseline > gpurun_out/r02_bench_cfg4_n1_samebox.json 2> gpurun_out/r02_bench_cfg4_n1_samebox.err
(timeout 300 python -m pytest tests/test_pipeline_pool.py -x -q > gpurun_out/r02_n2_pooltests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_n2_pooltests.log); tail -3 gpurun_out/r02_n2_pooltests.log
python - <<PY
import json
for f in ("r02_bench_cfg4_n2","r02_bench_cfg3_n2","r02_bench_cfg4_n1_samebox"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f,"n_gpus",d["n_gpus"],"value %.1f e2e %.1f ms/step %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"]))
    except Exception as ex: print(f,"failed",ex, open("gpurun_out/%s.err"%f).read()[-600:])
PY
