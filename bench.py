#!/usr/bin/env python
"""bench.py -- frames/s end-to-end (conv + PAF parse), OpenPose-COCO VGG-19 368x656, batch 16 per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU parser (oracle/_ref) on the host cores

One "step" = one pass of the hot path over one batch of 16 synthetic frames per GPU (weak scaling:
frames shard across GPUs): frame pre-processing + every conv of OpenPose-VGG19 (random-init
weights of the real architecture) + the PAF parse of the batch (+ for N>1 the NCCL gather of the
keypoint records to rank 0).  Because random weights give structureless heat-maps, seeded synthetic
crowd tensors are copied over the backbone's outputs after the last conv (hp_engine_set_output_override,
SURVEY 8d) -- all conv work is still executed; this is stated in config.parse_input.

  value : device-resident inputs (u8 frames already in HBM), results left on the device (rank 0 after gather)
  e2e   : the public host call hp_pose_run_u8_host -- pinned host frames H2D, human_t records D2H, every step
  roofline : the conv kernel (dominant): algorithmic FLOPs / CUDA-event time of the conv launches, measured in
             the timed region on the launching stream, vs the measured cuBLAS bf16 peak
  cpu_baseline : the reference's CPU parser timed on this host (parse stage only: the reference never runs
             the convs on a CPU; its engine is TensorRT)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# workloads (BASELINE.json configs).  The default -- and the only one the headline metric is quoted on -- is cfg3.
WORKLOADS = {
    "cfg3": dict(name="cfg3: OpenPose-COCO VGG-19 368x656, batch 16 per GPU, frame-sharded", graph="openpose_vgg19", in_h=368, in_w=656,
                 batch=16, persons=(10, 20), algo_flops=484.6e9, arch="OpenPose-VGG19 (6 stages), 484.6 GFLOP/frame"),
    "cfg2": dict(name="cfg2: Lightweight-OpenPose (MobilenetThin) 368x432, batch 8", graph="mobilenet_thin_openpose", in_h=368, in_w=432,
                 batch=8, persons=(1, 5), algo_flops=22.3e9, arch="MobilenetThin-OpenPose (6 stages), 22.3 GFLOP/frame"),
    "cfg4": dict(name="cfg4: OpenPose-ResNet50 (LW-OpenPose head, stride 8) 368x432, batch 32 per GPU, synthetic crowd", graph="resnet50_lw_openpose",
                 in_h=368, in_w=432, batch=32, persons=(10, 20), algo_flops=136.7e9, arch="ResNet50 + LW-OpenPose head, 136.7 GFLOP/frame"),
    "cfg5": dict(name="cfg5: OpenPifPaf ResNet50 385x385, batch 16 (pif/paf field decode)", graph="resnet50_pifpaf", in_h=385, in_w=385,
                 batch=16, persons=(2, 8), algo_flops=100.2e9, arch="ResNet50 (stride 16, no max-pool) + PIF/PAF heads, 100.2 GFLOP/frame", pifpaf=True),
}
IN_H, IN_W, HF, WF, BATCH, PERSONS, ALGO_FLOPS_PER_FRAME = 368, 656, 46, 82, 16, (10, 20), 484.6e9
WL = WORKLOADS["cfg3"]
N_INPUT_SETS = 12                    # distinct input batches rotated so that the frames alone exceed the 126 MB L2


def select_workload(key):
    global IN_H, IN_W, HF, WF, BATCH, PERSONS, ALGO_FLOPS_PER_FRAME, WL
    WL = WORKLOADS[key]
    IN_H, IN_W, BATCH, PERSONS, ALGO_FLOPS_PER_FRAME = WL["in_h"], WL["in_w"], WL["batch"], WL["persons"], WL["algo_flops"]
    HF, WF = IN_H // 8, IN_W // 8
    if WL.get("pifpaf"):
        HF, WF = (IN_H - 1) // 8 + 1, (IN_W - 1) // 8 + 1     # 49 x 49 fields for 385 x 385


def METRIC():
    return "frames/sec end-to-end (conv+PAF parse) OpenPose-COCO 368x656" if WL is WORKLOADS["cfg3"] else f"frames/sec end-to-end (conv+PAF parse) {WL['name']}"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d, "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)"""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def crowd_tensors(seed: int):
    from hyperpose_b200 import synthetic as syn
    return syn.make_batch_tensors(seed, BATCH, PERSONS, HF, WF)


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own src/paf.cpp (oracle/_ref) or the oracle port, parse stage only
# ------------------------------------------------------------------------------------------------
def cpu_parse_rate(conf, paf, seconds: float, threads: int):
    """frames/s of the CPU parser on the host cores; one parser replica per thread, frames round-robin
    (the reference's stream API does exactly this: stream.hpp:139,365-373)."""
    import oracle
    kind = "reference" if oracle.ref_available() else "port"
    n = conf.shape[0]
    counts = [0] * threads
    stop = time.time() + seconds

    def work(t):
        if kind == "reference":
            rp = oracle.RefParser()
            fn = lambda i: rp.process(conf[i], paf[i])
        else:
            fn = lambda i: oracle.oracle_process(conf[i], paf[i])
        i = t
        fn(i % n)  # first call allocates (paf.cpp:321-332): not timed
        t0 = time.time()
        while time.time() < stop:
            fn(i % n)
            i += threads
            counts[t] += 1
        return time.time() - t0

    t_start = time.time()
    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    wall = time.time() - t_start
    return sum(counts) / max(wall, 1e-9), kind


def cpu_conv_port(graph_name: str):
    """conv stage on the host cores: oracle/torch_backbone.py (plain PyTorch fp32, all cores) on ONE synthetic frame.
    The reference has no CPU implementation of its convs (TensorRT on a GPU, src/tensorrt.cpp:387-396), so this stage
    of the CPU arm is a port, the parse stage is the reference's own code."""
    import torch
    from hyperpose_b200 import models, synthetic as syn
    from oracle import torch_backbone
    graph = getattr(models, graph_name)(seed=0)
    frame = syn.make_frames_u8(2, 1, IN_H, IN_W)

    def run():
        with torch.no_grad():
            torch_backbone.run_graph(graph, frame, device="cpu")
    return run, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the path on this host's CPU cores, same metric (frames/s, conv + parse).  Each step is a bounded
    sample of the workload: ONE frame through the conv stage (PyTorch fp32 port on all cores -- the reference's convs are
    TensorRT-on-GPU and have no CPU implementation) and through the reference's own parser (src/paf.cpp via oracle/_ref)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    conf, paf = crowd_tensors(1000)
    cores = os.cpu_count() or 1
    kind = "reference" if oracle.ref_available() else "port"
    rp = oracle.RefParser() if kind == "reference" else None
    conv_one, conv_threads = cpu_conv_port(WL["graph"])
    t_conv = t_parse = 0.0

    def step(i, timed):
        nonlocal t_conv, t_parse
        t0 = time.time()
        conv_one()
        t1 = time.time()
        if kind == "reference":
            rp.process(conf[i % BATCH], paf[i % BATCH])
        else:
            oracle.oracle_process(conf[i % BATCH], paf[i % BATCH])
        t2 = time.time()
        if timed:
            t_conv += t1 - t0
            t_parse += t2 - t1

    for i in range(max(args.warmup, 1)):
        step(i, False)
    t0 = time.time()
    for i in range(args.steps):
        step(i, True)
    dt = time.time() - t0
    fps = args.steps / dt
    # the parser alone with every host thread (one replica per thread, stream.hpp:139): reported, not the value
    threads = min(cores, BATCH)
    parse_rate, _ = cpu_parse_rate(conf, paf, 3.0, threads)
    line = {
        "impl": "reference", "metric": METRIC(), "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WL["name"] + " -- CPU arm: one frame per step (bounded sample of the batch)",
                   "frames_per_step": 1, "persons_per_frame": list(PERSONS),
                   "conv_stage": f"PyTorch fp32 port of the same graph on {conv_threads} threads (the reference's convs are TensorRT-on-GPU: no CPU implementation exists)",
                   "parse_stage": "the reference's own src/paf.cpp compiled verbatim (oracle/_ref)" if kind == "reference" else "oracle port of src/paf.cpp"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": conv_threads, "kind": "port",
                         "sample": f"{args.steps} frames {IN_H}x{IN_W}: conv stage {t_conv / args.steps * 1e3:.0f} ms/frame (torch fp32 port, {conv_threads} threads) + "
                                   f"parse stage {t_parse / args.steps * 1e3:.1f} ms/frame ({kind} parser, 1 thread) of {cores} host cores",
                         "parse_only": {"value": parse_rate, "unit": "frames/s", "cores": threads, "kind": kind}},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from hyperpose_b200 import capi, models, synthetic as syn

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    graph = getattr(models, WL["graph"])(seed=0)
    pack = graph.to_pack()
    engine = capi.Engine(pack, (IN_W, IN_H), max_batch_size=BATCH, device=local_rank)
    del pack
    PIFPAF = bool(WL.get("pifpaf"))
    HCAP = 128 if PIFPAF else 64
    if PIFPAF:
        parser = capi.PifPafParser(IN_H, IN_W, 0.1, device=local_rank)
    else:
        parser = capi.PafParser(0.05, 0.05, device=local_rank)
        parser.set_capacity(peaks_per_part=128, candidates_per_limb=2048, humans=HCAP)

    # inputs: N_INPUT_SETS distinct batches of frames (device + pinned host), one set of crowd tensors per rank
    rng_seed = 2 + 1000 * rank
    frames_host = [torch.from_numpy(syn.make_frames_u8(rng_seed + i, BATCH, IN_H, IN_W)).pin_memory() for i in range(N_INPUT_SETS)]
    frames_dev = [f.to(dev) for f in frames_host]
    if PIFPAF:
        fields = [syn.make_pifpaf_fields(1000 * (rank + 1) + i, PERSONS, HF, WF) for i in range(BATCH)]
        conf_np = np.stack([f[0] for f in fields]).reshape(BATCH, 85, HF, WF)
        paf_np = np.stack([f[1] for f in fields]).reshape(BATCH, 171, HF, WF)
    else:
        conf_np, paf_np = crowd_tensors(1000 + rank)
    d_conf = torch.from_numpy(conf_np).to(dev)
    d_paf = torch.from_numpy(paf_np).to(dev)
    engine.set_output_override(d_conf.data_ptr(), d_paf.data_ptr())
    out_conf_ptr, out_paf_ptr, _ = engine.device_outputs()

    st = torch.cuda.Stream(device=dev)
    rec_bytes = capi.HUMAN_DT.itemsize
    # keypoint records + per-frame counts of one batch in ONE buffer (a single NCCL all-gather per step), double-buffered so
    # that the gather of batch i runs on a side stream while batch i+1 is computed
    hum_bytes = BATCH * HCAP * rec_bytes
    res_bufs = [torch.zeros(hum_bytes + BATCH * 4, dtype=torch.uint8, device=dev) for _ in range(2)]
    gath_bufs = [torch.zeros(world * (hum_bytes + BATCH * 4), dtype=torch.uint8, device=dev) for _ in range(2)] if world > 1 else None
    sg = torch.cuda.Stream(device=dev)
    ev_res = [torch.cuda.Event() for _ in range(2)]
    ev_gat = [torch.cuda.Event() for _ in range(2)]
    gstate = {"n": 0}
    def gather_results():
        if PIFPAF:
            return    # config 5 is a single-GPU config: records stay in the decoder's device buffer
        k = gstate["n"] & 1
        if gstate["n"] >= 2 and world > 1 and not os.environ.get("HPB_NO_GATHER"):
            st.wait_event(ev_gat[k])              # the gather that last read this buffer has finished
        buf = res_bufs[k]
        parser.copy_results_device(buf.data_ptr(), buf.data_ptr() + hum_bytes, BATCH, HCAP, st.cuda_stream)
        if world > 1 and not os.environ.get("HPB_NO_GATHER"):   # (diagnostic switch; the gather is part of the metric)
            ev_res[k].record(st)
            sg.wait_event(ev_res[k])
            with torch.cuda.stream(sg):
                dist.all_gather_into_tensor(gath_bufs[k], buf)   # ~300 KB per rank, off the conv stream
            ev_gat[k].record(sg)
        gstate["n"] += 1

    def drain_gather():
        if world > 1 and not PIFPAF and not os.environ.get("HPB_NO_GATHER"):
            for k in range(2):
                if gstate["n"] > k:
                    st.wait_event(ev_gat[k])

    # Optional (--pipeline; OFF by default: measured SLOWER, 2070 vs 2221 frames/s -- the parser's many small CTAs delay
    # the start of the next persistent conv kernel's CTAs, whose static tile assignment then runs unbalanced).
    # Software pipeline of the device-resident path: the PAF parse (+ the keypoint gather) of batch i runs on a second
    # stream while the convs of batch i+1 run on the first.  The engine's conf/paf outputs are snapshotted (D2D, 14 MB)
    # on the conv stream so that batch i+1 may overwrite them; events order snapshot <-> parse in both directions.
    st2 = torch.cuda.Stream(device=dev)
    snap_conf = torch.empty(BATCH * 19 * HF * WF, dtype=torch.float32, device=dev)
    snap_paf = torch.empty(BATCH * 38 * HF * WF, dtype=torch.float32, device=dev)
    ev_ready = torch.cuda.Event()
    ev_parsed = torch.cuda.Event()
    state = {"first": True}

    def step_device_serial(i):
        engine.infer_u8_device(frames_dev[i % N_INPUT_SETS].data_ptr(), BATCH, st.cuda_stream)
        if PIFPAF:
            parser.process_device(out_conf_ptr, out_paf_ptr, BATCH, HF, WF, st.cuda_stream)
        else:
            parser.process_device(out_conf_ptr, out_paf_ptr, BATCH, 19, 38, HF, WF, st.cuda_stream)
        gather_results()

    def step_device(i):
        if not args.pipeline:
            return step_device_serial(i)
        engine.infer_u8_device(frames_dev[i % N_INPUT_SETS].data_ptr(), BATCH, st.cuda_stream)
        if not state["first"]:
            st.wait_event(ev_parsed)                  # the previous parse has finished reading the snapshot
        state["first"] = False
        engine.copy_outputs_device(snap_conf.data_ptr(), snap_paf.data_ptr(), BATCH, st.cuda_stream)
        ev_ready.record(st)
        st2.wait_event(ev_ready)
        parser.process_device(snap_conf.data_ptr(), snap_paf.data_ptr(), BATCH, 19, 38, HF, WF, st2.cuda_stream)
        parser.copy_results_device(res_bufs[0].data_ptr(), res_bufs[0].data_ptr() + hum_bytes, BATCH, HCAP, st2.cuda_stream)
        if world > 1:
            with torch.cuda.stream(st2):
                dist.all_gather_into_tensor(gath_bufs[0], res_bufs[0])
        ev_parsed.record(st2)

    def drain_device():
        if args.pipeline:
            st.wait_event(ev_parsed)                      # the timed region ends when the last parse/gather has finished

    def step_host(i):
        if PIFPAF:   # engine.inference(batch) + pifpaf.process per image: host frames in, humans out (fields stay on the device)
            engine.infer_u8(frames_host[i % N_INPUT_SETS].numpy())
            _, _, es = engine.device_outputs()
            parser.process_device(out_conf_ptr, out_paf_ptr, BATCH, HF, WF, es)
            return parser.fetch(BATCH, cap=HCAP)
        humans = engine.run_pose(parser, frames_host[i % N_INPUT_SETS].numpy(), cap=HCAP)
        if world > 1:
            gather_results()
            sg.synchronize()
        return humans

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False):
        for i in range(warmup):
            fn(i)
        barrier()
        l0 = engine.launch_count + parser.launch_count
        if profile:
            engine.set_profiling(True)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        t0 = time.time()
        for i in range(steps):
            fn(warmup + i)
        if fn is step_device:
            drain_device()
        drain_gather()                                # the timed region ends when the last keypoint gather has finished
        e1.record(st)
        torch.cuda.synchronize()
        wall = time.time() - t0
        ms_dev = e0.elapsed_time(e1)
        if profile:
            engine.set_profiling(False)
        launches = engine.launch_count + parser.launch_count - l0
        barrier()
        ms = max(ms_dev, 0.0)
        # host-synchronous paths are bounded by wall clock, device-async ones by the stream events: take the larger
        ms = max(ms, wall * 1e3) if fn is step_host else ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches

    # sanity: the device path and the host path agree with each other before anything is timed
    step_device(0)
    torch.cuda.synchronize()
    state["first"] = True
    ref_h = parser.fetch(BATCH, cap=HCAP)
    host_h = step_host(0)
    assert all(a.tobytes() == b.tobytes() for a, b in zip(ref_h, host_h)), "device and host paths disagree"
    n_humans = sum(len(h) for h in host_h)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(30):                     # (every rank: the steps contain the collective) nvidia-smi needs ~0.3 s to deliver
        step_device(i)                      # its first sample: keep the GPUs under the same load meanwhile
    torch.cuda.synchronize()
    sampler.lines.clear()                   # keep only samples taken under load
    ms_total, launches = timed(step_device, args.steps, args.warmup, profile=True)
    clocks = sampler.stop() if rank == 0 else None
    prof_ms, prof_ty, prof_fl, prof_runs = engine.get_profile()
    ms_e2e, _ = timed(step_host, max(3, args.steps // 2), max(3, args.warmup))
    e2e_steps = max(3, args.steps // 2)

    frames_total = world * BATCH * args.steps
    value = frames_total / (ms_total / 1e3)
    e2e_value = world * BATCH * e2e_steps / (ms_e2e / 1e3)

    if rank == 0:
        peaks, peak_src = load_peaks()
        conv_ms = float(prof_ms[prof_ty == models.OP_CONV].sum())
        other_ms = float(prof_ms[prof_ty != models.OP_CONV].sum())
        n_conv = int((prof_ty == models.OP_CONV).sum())
        algo_flops = ALGO_FLOPS_PER_FRAME * BATCH
        achieved = algo_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")
        if os.path.exists(tpath) and args.workload == "cfg3":      # the ncu capture is of the headline workload only
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_step")
            except Exception:
                traffic = None
        # bounded CPU baseline (rank 0, N=1 only): ~12 s of the reference parser on the host cores
        cpu = None
        if world == 1 and not args.no_cpu_baseline and not PIFPAF:
            cores = os.cpu_count() or 1
            threads = min(cores, BATCH)
            rate, kind = cpu_parse_rate(conf_np, paf_np, 12.0, threads)
            conv_one, conv_threads = cpu_conv_port(WL["graph"])
            conv_one()                                   # warm-up (allocations, oneDNN primitive caches)
            tc0 = time.time(); n_cpu_frames = 0
            while n_cpu_frames < 2 or (time.time() - tc0 < 6.0 and n_cpu_frames < 16):
                conv_one(); n_cpu_frames += 1
            conv_s = (time.time() - tc0) / n_cpu_frames
            whole = 1.0 / (conv_s + 1.0 / rate)
            cpu = {"value": whole, "unit": "frames/s", "cores": max(threads, conv_threads), "kind": "port",
                   "sample": f"conv stage: {n_cpu_frames} frames {IN_H}x{IN_W} through a PyTorch fp32 port of the same graph on {conv_threads} threads ({conv_s * 1e3:.0f} ms/frame; the reference's "
                             f"convs are TensorRT-on-GPU, no CPU implementation exists) + parse stage: 12 s of the reference CPU parser (src/paf.cpp via oracle/_ref) on the step's {BATCH} synthetic "
                             f"{HF}x{WF} crowd frames, {threads} threads of {cores} host cores",
                   "parse_only": {"value": rate, "unit": "frames/s", "cores": threads, "kind": kind}}
        layers = [{"op": i, "name": graph.ops[i].name, "type": int(prof_ty[i]), "ms": float(prof_ms[i]),
                   "tflops": (float(prof_fl[i]) * BATCH / (prof_ms[i] / 1e3) / 1e12 if prof_ms[i] > 0 and prof_fl[i] > 0 else None)}
                  for i in range(len(prof_ms))]
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"bench_layers_n{world}.json"), "w") as f:
            json.dump({"ms_per_step": ms_total / args.steps, "conv_ms": conv_ms, "other_engine_ms": other_ms, "profiled_runs": prof_runs, "layers": layers}, f, indent=1)
        h2d = BATCH * IN_H * IN_W * 3
        d2h = BATCH * HCAP * rec_bytes + BATCH * 8
        line = {
            "metric": METRIC(), "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate (conv); f32/f64 (parse)", "data": "synthetic",
            "config": {"workload": WL["name"],
                       "global_batch": world * BATCH, "input": f"u8 frames {IN_H}x{IN_W}x3, random (default_rng)",
                       "weights": "random-init (He-normal, seed 0) of the reference architecture: " + WL["arch"],
                       "parse_input": f"synthetic {'PIF/PAF fields' if PIFPAF else 'crowd tensors'} ({PERSONS[0]}-{PERSONS[1]} persons/frame, {n_humans} humans/batch) copied over the conv outputs after the last conv",
                       "l2": f"{N_INPUT_SETS} distinct input batches rotated ({N_INPUT_SETS * BATCH * IN_H * IN_W * 3 / 1e6:.0f} MB > L2); activations (>1 GB/step) stream through",
                       "parallelism": f"dp{world} (frames shard; NCCL all-gather of keypoint records only)" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "hp_pose_run_u8_host (pinned host frames in, human_t records out, synchronous per batch)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                         "traffic": traffic, "kernel": "conv_tcgen05 family (conv_tcgen05_kernel / _swap_kernel / conv_halo_kernel / conv_stem3_kernel)", "launches_per_step": n_conv,
                         "algorithmic_flops_per_step": algo_flops, "kernel_ms_per_step": conv_ms,
                         "kernel_share_of_step": conv_ms / (ms_total / args.steps), "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)"},
            "clocks": clocks,
        }
        # backbone-only and parser-only rates of the same run (per GPU), and the parser against its HBM bound
        step_ms = ms_total / args.steps
        parse_ms = max(step_ms - conv_ms - other_ms, 1e-6)
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        # (19+38)*Hf*Wf*4 B per frame for conf/PAF; (17*5+19*9)*h*w*4 B for the PIF/PAF fields
        parse_bytes = BATCH * ((17 * 5 + 19 * 9) if PIFPAF else 57) * HF * WF * 4
        line["breakdown"] = {"backbone_ms_per_step": conv_ms + other_ms, "backbone_frames_per_s": BATCH / ((conv_ms + other_ms) / 1e3),
                             "parse_ms_per_step": parse_ms, "parse_frames_per_s": BATCH / (parse_ms / 1e3),
                             "parse_hbm_frac": (parse_bytes / (parse_ms / 1e3) / 1e9 / hbm) if parse_bytes else None,
                             "parse_algorithmic_bytes_per_step": parse_bytes,
                             "note": "parse = step minus the engine's per-op events (parser kernels + result copy); its HBM bound counts only the network-output tensors read once"}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    engine.close()
    parser.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline", action="store_true", help="two-stream software pipeline (parse of batch i overlaps convs of batch i+1)")
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS), help="BASELINE.json config (default: the headline cfg3)")
    args = ap.parse_args()
    select_workload(args.workload)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
