"""Diagnostic: per-op event times of the first ops of the cfg3 graph in three contexts
(engine only / engine + parser on the same stream / rotating input buffers)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperpose_b200 import capi, models, synthetic as syn  # noqa: E402

H, W, HF, WF, B = 368, 656, 46, 82, 16
eng = capi.Engine(models.openpose_vgg19(0).to_pack(), (W, H), max_batch_size=B)
parser = capi.PafParser()
parser.set_capacity(128, 2048, 64)
sets = [torch.from_numpy(syn.make_frames_u8(2 + i, B, H, W)).cuda() for i in range(12)]
conf, paf = syn.make_batch_tensors(1000, B, (10, 20), HF, WF)
dc, dp = torch.from_numpy(conf).cuda(), torch.from_numpy(paf).cuda()
eng.set_output_override(dc.data_ptr(), dp.data_ptr())
oc, op, _ = eng.device_outputs()
st = torch.cuda.Stream()


def run(label, steps, with_parser, rotate, sync_each):
    eng.set_profiling(False)
    for i in range(3):
        eng.infer_u8_device(sets[0].data_ptr(), B, st.cuda_stream)
    torch.cuda.synchronize()
    eng.set_profiling(True)
    t0 = time.time()
    for i in range(steps):
        eng.infer_u8_device(sets[i % 12 if rotate else 0].data_ptr(), B, st.cuda_stream)
        if with_parser:
            parser.process_device(oc, op, B, 19, 38, HF, WF, st.cuda_stream)
        if sync_each:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    dt = (time.time() - t0) / steps * 1e3
    ms, ty, fl, runs = eng.get_profile()
    print(f"{label:40s} step {dt:6.3f} ms | im2col {ms[0]*1e3:6.1f} conv1_1 {ms[1]*1e3:6.1f} conv1_2 {ms[2]*1e3:6.1f} pool {ms[3]*1e3:6.1f} conv2_1 {ms[4]*1e3:6.1f} | sum {ms.sum():6.3f} runs {runs}")


for steps in (10, 60):
    run(f"engine only, {steps} steps", steps, False, False, False)
    run(f"engine+parser, {steps} steps", steps, True, False, False)
    run(f"engine+parser rotating, {steps} steps", steps, True, True, False)
    run(f"engine+parser rotating sync, {steps} steps", steps, True, True, True)
