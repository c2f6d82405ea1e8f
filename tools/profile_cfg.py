"""one forward of a workload's network (no parser), for ncu:  ncu ... python tools/profile_cfg.py --graph mobilenet_thin_openpose --h 368 --w 432 --batch 8"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperpose_b200 import capi, models, synthetic as syn
ap = argparse.ArgumentParser()
ap.add_argument("--graph", default="mobilenet_thin_openpose"); ap.add_argument("--h", type=int, default=368); ap.add_argument("--w", type=int, default=432)
ap.add_argument("--batch", type=int, default=8); ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
eng = capi.Engine(getattr(models, a.graph)(0).to_pack(), (a.w, a.h), max_batch_size=a.batch)
fr = syn.make_frames_u8(1, a.batch, a.h, a.w)
for _ in range(a.steps):
    eng.infer_u8(fr)
eng.sync()
print("ok")
