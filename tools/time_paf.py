"""scratch timing of the parser kernels (device-resident inputs); not the bench."""
import sys, time
import numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from hyperpose_b200 import capi, synthetic as syn
for (hf, wf, P, N) in [(46, 82, (3, 9), 16), (46, 54, (10, 20), 32), (46, 82, (3, 9), 64)]:
    conf, paf = syn.make_batch_tensors(21, N, P, hf, wf)
    dc = torch.from_numpy(conf).cuda(); dp = torch.from_numpy(paf).cuda()
    parser = capi.PafParser()
    st = torch.cuda.Stream()
    for _ in range(3):
        parser.process_device(dc.data_ptr(), dp.data_ptr(), N, 19, 38, hf, wf, st.cuda_stream)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(20):
            parser.process_device(dc.data_ptr(), dp.data_ptr(), N, 19, 38, hf, wf, st.cuda_stream)
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    got = parser.fetch(N)
    t = time.time(); got2 = parser.process_batch(conf, paf); host_ms = (time.time() - t) * 1e3
    print(f"{hf}x{wf} N={N}: device {ms*1e3:.1f} us/batch = {N/ms*1e3:.0f} frames/s; host-API {host_ms:.2f} ms/batch; humans {sum(len(g) for g in got)}")
