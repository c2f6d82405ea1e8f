"""Runs `--steps` device-resident steps of the bench workload (cfg3, batch 16) with nothing else around them,
for ncu captures:  ncu ... python tools/profile_step.py --steps 2
(step 0 is the warm-up; per step: 1 D2D memcpy, im2col, 52 conv, 3 maxpool, 2 D2D, memset + 2 parser kernels)"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperpose_b200 import capi, models, synthetic as syn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=16)
a = ap.parse_args()
H, W, HF, WF = 368, 656, 46, 82
eng = capi.Engine(models.openpose_vgg19(0).to_pack(), (W, H), max_batch_size=a.batch)
parser = capi.PafParser()
parser.set_capacity(128, 2048, 64)
frames = torch.from_numpy(syn.make_frames_u8(2, a.batch, H, W)).cuda()
conf, paf = syn.make_batch_tensors(1000, a.batch, (10, 20), HF, WF)
dc, dp = torch.from_numpy(conf).cuda(), torch.from_numpy(paf).cuda()
eng.set_output_override(dc.data_ptr(), dp.data_ptr())
oc, op, _ = eng.device_outputs()
st = torch.cuda.Stream()
torch.cuda.synchronize()
for i in range(a.steps):
    eng.infer_u8_device(frames.data_ptr(), a.batch, st.cuda_stream)
    parser.process_device(oc, op, a.batch, 19, 38, HF, WF, st.cuda_stream)
torch.cuda.synchronize()
print("humans", sum(len(h) for h in parser.fetch(a.batch, 64)))
