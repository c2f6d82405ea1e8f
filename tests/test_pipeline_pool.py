"""The pipelined end-to-end call (hp_pose_submit_u8_host / hp_pose_collect: two batches in flight, CUDA-graph replay) and
the in-process multi-GPU pool (hp_pool_*: one host thread + engine + parser per GPU, frames sharded in blocks of max_batch,
SURVEY 8e).  Everything goes through the C ABI; results are compared byte-for-byte with the oracle parser run on the tensors
the engine itself produced (parse parity is defined on identical tensors) and with the synchronous call."""
import numpy as np
import pytest

import oracle
from hyperpose_b200 import capi, models, synthetic as syn

pytestmark = pytest.mark.gpu

H, W = 64, 96


def _setup(N, seed=4):
    g = models.tiny_test_net(seed)
    eng = capi.Engine(g.to_pack(), (W, H), max_batch_size=N)
    return g, eng


def _thresholds(eng, frames):
    eng.infer_u8(frames)
    conf, paf = eng.read_outputs(frames.shape[0])
    return float(np.quantile(conf[:, :18], 0.97)), float(np.quantile(paf, 0.5))


_PEAKS = {"n": 0}


def _oracle_humans(eng, frames, ct, pt):
    eng.infer_u8(frames)
    conf, paf = eng.read_outputs(frames.shape[0])
    out = []
    for i in range(frames.shape[0]):
        o = oracle.oracle_process(conf[i], paf[i], ct, pt, peak_cap=1 << 18, conn_cap=1 << 14)
        _PEAKS["n"] += len(o["peaks"])
        out.append(o["humans"].tobytes())
    return out


def test_submit_collect_two_in_flight_equals_synchronous_call_and_oracle():
    N = 4
    g, eng = _setup(N)
    batches = [syn.make_frames_u8(40 + k, N, H, W) for k in range(5)]
    batches[3] = batches[3][:2]                                   # a smaller batch in the middle: the graph is re-captured
    ct, pt = _thresholds(eng, batches[0])
    parser = capi.PafParser(ct, pt)
    parser.set_capacity(peaks_per_part=1024, candidates_per_limb=1 << 15, humans=128)
    want = [_oracle_humans(eng, b, ct, pt) for b in batches]
    # software pipeline: submit k+1, then collect k
    got = [None] * len(batches)
    t_prev = eng.submit_pose(parser, batches[0])
    for k in range(1, len(batches)):
        t = eng.submit_pose(parser, batches[k])
        assert t != t_prev
        got[k - 1] = eng.collect_pose(t_prev, cap=128)
        t_prev = t
    got[-1] = eng.collect_pose(t_prev, cap=128)
    for k, b in enumerate(batches):
        assert [h.tobytes() for h in got[k]] == want[k], f"batch {k}"
        sync = eng.run_pose(parser, b, cap=128)
        assert [h.tobytes() for h in sync] == want[k]
    assert _PEAKS["n"] > 50, "vacuous: no peaks at this threshold"
    st = eng.pose_stats()
    assert st["graph_launches"] >= len(batches), st              # replayed from CUDA graphs, not launched one by one
    assert 2 <= st["graph_captures"] <= 8, st
    eng.close(); parser.close()


def test_third_submit_without_collect_is_refused_and_tickets_are_checked():
    N = 2
    g, eng = _setup(N)
    parser = capi.PafParser()
    fr = syn.make_frames_u8(1, N, H, W)
    t0 = eng.submit_pose(parser, fr)
    t1 = eng.submit_pose(parser, fr)
    with pytest.raises(capi.HyperposeError) as e:
        eng.submit_pose(parser, fr)
    assert e.value.status == capi.HP_ERR_ARG
    eng.collect_pose(t0); eng.collect_pose(t1)
    with pytest.raises(capi.HyperposeError):
        eng.collect_pose(t0)                                      # not in flight any more
    with pytest.raises(capi.HyperposeError) as e:
        eng.submit_pose(parser, np.zeros((N + 1, H, W, 3), np.uint8))
    assert e.value.status == capi.HP_ERR_BATCH                    # std::logic_error in the reference (tensorrt.cpp:439-443)
    eng.close(); parser.close()


def test_parser_capacity_overflow_grows_and_reruns_like_the_unbounded_reference():
    N = 3
    g, eng = _setup(N)
    fr = syn.make_frames_u8(7, N, H, W)
    ct, pt = _thresholds(eng, fr)
    want = _oracle_humans(eng, fr, ct, pt)
    parser = capi.PafParser(ct, pt)
    parser.set_capacity(peaks_per_part=2, candidates_per_limb=2, humans=1)     # everything overflows
    got = eng.run_pose(parser, fr, cap=256)
    assert [h.tobytes() for h in got] == want
    # and again, pipelined, with the grown capacities baked into a fresh graph
    t = eng.submit_pose(parser, fr)
    assert [h.tobytes() for h in eng.collect_pose(t, cap=256)] == want
    eng.close(); parser.close()


def test_malformed_packs_are_rejected():
    import struct
    pack = bytearray(models.tiny_test_net(0).to_pack())
    hdr = struct.calcsize("<8s6I3f5IQ")
    n_buffers, n_ops = struct.unpack_from("<2I", pack, 12)
    op0 = hdr + 8 * n_buffers
    # (a) truncated
    with pytest.raises(capi.HyperposeError) as e:
        capi.Engine(bytes(pack[: len(pack) // 2]), (W, H), max_batch_size=1)
    assert e.value.status == capi.HP_ERR_ARG
    # (b) a conv whose weight offset points past the blob
    for i in range(n_ops):
        if struct.unpack_from("<I", pack, op0 + 96 * i)[0] == models.OP_CONV:
            bad = bytearray(pack)
            struct.pack_into("<Q", bad, op0 + 96 * i + 72, 1 << 40)
            with pytest.raises(capi.HyperposeError) as e:
                capi.Engine(bytes(bad), (W, H), max_batch_size=1)
            assert e.value.status == capi.HP_ERR_ARG
            bad = bytearray(pack)
            struct.pack_into("<I", bad, op0 + 96 * i + 28, 0)     # groups = 0
            with pytest.raises(capi.HyperposeError) as e:
                capi.Engine(bytes(bad), (W, H), max_batch_size=1)
            assert e.value.status == capi.HP_ERR_ARG
            break
    else:
        pytest.fail("no conv op found")
    # (c) absurd counts in the header
    bad = bytearray(pack)
    struct.pack_into("<I", bad, 16, 0x7fffffff)
    with pytest.raises(capi.HyperposeError) as e:
        capi.Engine(bytes(bad), (W, H), max_batch_size=1)
    assert e.value.status == capi.HP_ERR_ARG


@pytest.mark.parametrize("n_frames", [13, 4, 1])
def test_pool_world_size_2_shards_blocks_and_returns_frame_order(n_frames):
    """world size 2 through the pool, no torch.distributed involved: two workers (GPUs 0 and 1 when the box has two, else two
    workers on GPU 0), max_batch 3 => blocks [0,3) [3,6) ... alternate between the workers; the humans of every frame equal
    the single-engine result for that frame, in frame order"""
    B = 3
    g = models.tiny_test_net(4)
    ndev = capi.lib().hp_device_count()
    devices = [0, 1] if ndev >= 2 else [0, 0]
    frames = syn.make_frames_u8(90, n_frames, H, W)
    eng = capi.Engine(g.to_pack(), (W, H), max_batch_size=B)
    ct, pt = _thresholds(eng, frames[:B])
    want = []
    for k in range(0, n_frames, B):
        want += _oracle_humans(eng, frames[k:k + B], ct, pt)
    eng.close()
    pool = capi.Pool(g.to_pack(), (W, H), B, devices=devices, conf_thresh=ct, paf_thresh=pt)
    pool.set_capacity(peaks_per_part=1024, candidates_per_limb=1 << 15, humans=128)
    for _ in range(2):                                            # second call: worker threads and graphs are reused
        got = pool.run(frames, cap=128)
        assert [h.tobytes() for h in got] == want
    assert pool.launch_count > 0
    pool.close()


def test_default_device_selector(monkeypatch):
    L = capi.lib()
    capi.handoff_stats()                                          # binds the engine-side prototypes
    monkeypatch.delenv("HPB_DEVICE", raising=False)
    assert L.hp_default_device() == 0
    n = L.hp_device_count()
    monkeypatch.setenv("HPB_DEVICE", "rr")
    seq = [L.hp_default_device() for _ in range(2 * n)]
    assert sorted(set(seq)) == list(range(n))
    monkeypatch.setenv("HPB_DEVICE", str(n - 1))
    assert L.hp_default_device() == n - 1
    monkeypatch.setenv("HPB_DEVICE", "99")
    assert L.hp_default_device() == 0


def test_pipelined_pifpaf_call_equals_the_plain_sequence():
    """hp_pose_submit_pifpaf_u8_host / hp_pose_collect (decoder on its own stream underneath the next batch's convolutions, SMs
    reserved for it) against engine.inference + decoder.process of the same fields; synthetic PIF / PAF fields are copied over the
    random-weight network's outputs (hp_engine_set_output_override) so that there are people to find."""
    import torch
    N, HW = 2, 129
    eng = capi.Engine(models.resnet50_pifpaf(0).to_pack(), (HW, HW), max_batch_size=N)
    hf = wf = (HW - 1) // 8 + 1
    fields = [syn.make_pifpaf_fields(70 + i, (1, 3), hf, wf) for i in range(N)]
    pif = np.stack([f[0] for f in fields]).astype(np.float32); paf = np.stack([f[1] for f in fields]).astype(np.float32)
    d_pif, d_paf = torch.from_numpy(pif).cuda(), torch.from_numpy(paf).cuda()
    eng.set_output_override(d_pif.data_ptr(), d_paf.data_ptr())
    dec = capi.PifPafParser(HW, HW, 0.1)
    want = [h.tobytes() for h in dec.process_batch(pif, paf)]
    assert sum(len(w) for w in want) > 0
    batches = [syn.make_frames_u8(90 + k, N, HW, HW) for k in range(4)]
    got = []
    t_prev = eng.submit_pose(dec, batches[0])
    for k in range(1, len(batches)):
        t = eng.submit_pose(dec, batches[k])
        assert t != t_prev
        got.append(eng.collect_pose(t_prev, cap=128))
        t_prev = t
    got.append(eng.collect_pose(t_prev, cap=128))
    for g in got:
        assert [h.tobytes() for h in g] == want
    # device-resident frames, one batch in flight
    d_frames = torch.from_numpy(batches[0]).cuda()
    g = eng.collect_pose(eng.submit_pose_device(dec, d_frames.data_ptr(), N), cap=128)
    assert [h.tobytes() for h in g] == want
    # a PAF-parser handle is refused on an OpenPifPaf pack and vice versa
    with pytest.raises(capi.HyperposeError):
        eng.submit_pose(capi.PafParser(), batches[0])
    eng.close(); dec.close()
