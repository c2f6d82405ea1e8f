"""Small invocations of every kernel family for compute-sanitizer:
   compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py
(tiny conv network incl. all op kinds, PAF parser, PifPaf decoder, Pose Proposal parser, frame resize)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperpose_b200 import capi, models, synthetic as syn  # noqa: E402

which = sys.argv[1:] or ["engine", "paf", "pifpaf", "ppn"]
if "engine" in which:
    H, W, N = 64, 96, 2
    eng = capi.Engine(models.tiny_test_net(4).to_pack(), (W, H), max_batch_size=N)
    frames = syn.make_frames_u8(9, N, H, W)
    eng.infer_u8(frames)
    conf, paf = eng.read_outputs(N)
    big = syn.make_frames_u8(3, 1, 150, 201)[0]
    eng.stage_frame(0, big, keep_ratio=True); eng.stage_frame(1, big, keep_ratio=False)
    eng.infer_staged(N)
    eng.sync()
    print("engine ok", conf.shape, float(np.abs(conf).max()))
    eng.close()
if "paf" in which:
    conf, paf = syn.make_batch_tensors(0, 2, (2, 4), 46, 54)
    p = capi.PafParser()
    print("paf ok", [len(h) for h in p.process_batch(conf, paf)])
    p.close()
if "pifpaf" in which:
    pif, paf = syn.make_pifpaf_fields(10, 2, 25, 33)
    d = capi.PifPafParser(193, 257, 0.1)
    print("pifpaf ok", len(d.process(pif, paf)))
    d.close()
if "ppn" in which:
    t = syn.make_ppn_tensors(21, 4)
    q = capi.PoseProposalParser((384, 384))
    print("ppn ok", len(q.process(*t)))
    q.set_point_thresh(0.05); q.set_limb_thresh(0.03)
    print("ppn dense ok", len(q.process(*syn.make_ppn_tensors(26, 5, distractors=60))))
    q.close()
