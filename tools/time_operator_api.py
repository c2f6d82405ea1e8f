"""Operator-API-shaped timing (examples/operator_api_batched_images_paf.example.cpp:64-74): per batch
    engine.inference(host frames) -> per-image host feature maps -> parser.process(conf_i, paf_i) per image
with the device-resident hand-off (csrc/handoff.h) on and off, next to the fused C-ABI call hp_pose_run_u8_host.
cfg3 geometry (OpenPose-VGG19 368x656, batch 16), synthetic crowd tensors over the conv outputs (bench hook)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperpose_b200 import capi, models, synthetic as syn  # noqa: E402

H, W, HF, WF, B = 368, 656, 46, 82, 16
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
eng = capi.Engine(models.openpose_vgg19(0).to_pack(), (W, H), max_batch_size=B)
parser = capi.PafParser()
parser.set_capacity(128, 2048, 64)
frames = syn.make_frames_u8(2, B, H, W)
conf, paf = syn.make_batch_tensors(1000, B, (10, 20), HF, WF)
dc, dp = torch.from_numpy(conf).cuda(), torch.from_numpy(paf).cuda()
torch.cuda.synchronize()
eng.set_output_override(dc.data_ptr(), dp.data_ptr())


def operator_api(n):
    humans = 0
    for _ in range(n):
        eng.infer_u8(frames)
        packets = eng.read_outputs_frames(B)
        for c, p in packets:
            humans += len(parser.process(c, p))
    return humans


def fused(n):
    humans = 0
    for _ in range(n):
        humans += sum(len(h) for h in eng.run_pose(parser, frames, cap=64))
    return humans


out = {}
for label, fn, on in (("operator_api_handoff", operator_api, True), ("operator_api_host_roundtrip", operator_api, False), ("fused_c_abi", fused, True)):
    capi.handoff_enable(on)
    fn(3)
    t0 = time.perf_counter()
    h = fn(steps)
    dt = time.perf_counter() - t0
    out[label] = {"frames_per_s": steps * B / dt, "ms_per_batch": dt / steps * 1e3, "humans": h}
    print(label, out[label], flush=True)
capi.handoff_enable(True)
print("handoff", capi.handoff_stats())
import json
print(json.dumps(out))
