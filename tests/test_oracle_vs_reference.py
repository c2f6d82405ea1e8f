"""The oracle restatement (oracle/paf_oracle.c) against the reference's OWN parser:
  * committed goldens: human_t lists that /root/reference/src/paf.cpp (compiled verbatim
    into oracle/_ref) produced on the seeded synthetic frames -- runs anywhere;
  * live oracle/_ref, when it is built (build container, or prebuilt .so on the GPU box)."""
import os

import numpy as np
import pytest

import oracle
from hyperpose_b200 import synthetic as syn
from tests.golden.make_golden import AREA_FRAME_CASES, FRAME_CASES, sha


@pytest.fixture(scope="module")
def ref_humans(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_humans.npz"))


@pytest.mark.parametrize("case", FRAME_CASES, ids=[c[0] for c in FRAME_CASES])
def test_oracle_matches_reference_golden(ref_humans, case):
    name, seed, P, hf, wf, rw, rh, ct, pt = case
    conf, paf = syn.make_frame_tensors(seed, P, hf, wf)
    assert sha(conf) + sha(paf) == str(ref_humans[name + "_in_sha"]), "synthetic generator drifted from the goldens"
    got = oracle.oracle_process(conf, paf, ct, pt, rw, rh)["humans"]
    want = ref_humans[name + "_humans"]
    assert len(got) == len(want)
    assert got.tobytes() == want.tobytes()      # byte-identical human_t records, same order


@pytest.mark.parametrize("case", AREA_FRAME_CASES, ids=[c[0] for c in AREA_FRAME_CASES])
def test_oracle_matches_reference_golden_on_shrinking_resolutions(golden_dir, case):
    """A5 beyond pure up-scaling: resolutions that shrink an axis (mixed 2-tap regime, integer and fractional area averaging).
    Goldens = the reference's own src/paf.cpp over the cv2-pinned resize (tests/golden/make_golden.py area)."""
    ref = np.load(os.path.join(golden_dir, "ref_humans_area.npz"))
    name, seed, P, hf, wf, rw, rh, ct, pt = case
    conf, paf = syn.make_frame_tensors(seed, P, hf, wf)
    assert sha(conf) + sha(paf) == str(ref[name + "_in_sha"]), "synthetic generator drifted from the goldens"
    o = oracle.oracle_process(conf, paf, ct, pt, rw, rh)
    assert o["humans"].tobytes() == ref[name + "_humans"].tobytes()
    if oracle.ref_available():
        rp = oracle.RefParser(ct, pt, rw, rh)
        assert rp.process(conf, paf).tobytes() == o["humans"].tobytes()
        rp.close()


def test_goldens_are_not_vacuous(ref_humans):
    assert len(ref_humans["cfg1_p1_humans"]) == 1
    assert len(ref_humans["cfg1_p3_humans"]) >= 2
    assert len(ref_humans["cfg4_crowd_humans"]) >= 10
    assert len(ref_humans["empty_humans"]) == 0
    h = ref_humans["cfg1_p1_humans"][0]
    assert int(h["parts"]["has_value"].sum()) == 18


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", range(20, 32))
def test_oracle_matches_live_reference_random(seed):
    rng = np.random.default_rng(seed)
    hf, wf = [(46, 54), (46, 82), (32, 32), (23, 40)][seed % 4]
    P = int(rng.integers(0, 16))
    conf, paf = syn.make_frame_tensors(seed, P, hf, wf)
    rp = oracle.RefParser()
    want = rp.process(conf, paf)
    rp.close()
    got = oracle.oracle_process(conf, paf)["humans"]
    assert got.tobytes() == want.tobytes()


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built")
def test_oracle_matches_live_reference_noise_field():
    # structureless input (many spurious peaks, few/no humans): exercises the candidate filter
    rng = np.random.default_rng(99)
    conf = rng.random((19, 24, 30), dtype=np.float32) * 0.3
    paf = (rng.random((38, 24, 30), dtype=np.float32) - 0.5)
    rp = oracle.RefParser()
    want = rp.process(conf, paf)
    rp.close()
    o = oracle.oracle_process(conf, paf)
    assert len(o["peaks"]) > 50
    assert o["humans"].tobytes() == want.tobytes()


def test_peak_ids_are_scan_ordered():
    conf, paf = syn.make_frame_tensors(3, (10, 20), 46, 54)
    pk = oracle.oracle_process(conf, paf)["peaks"]
    assert np.array_equal(pk["id"], np.arange(len(pk)))
    key = pk["part_id"].astype(np.int64) * (1 << 32) + pk["y"].astype(np.int64) * (1 << 16) + pk["x"]
    assert np.all(np.diff(key) > 0)


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", range(100, 140))
def test_oracle_matches_live_reference_sweep(seed):
    """wider sweep of what pins the oracle: odd feature sizes, both thresholds, custom parser resolutions (non-integer
    up-scaling factors -> the 2-tap INTER_AREA path), empty frames and crowds, noise levels around the thresholds.
    Two regions are left out because the reference itself leaves them open, not because the restatement differs:
      * noise-free fields: their plateaus give limb candidates with EXACTLY equal scores, and paf.cpp:246 orders
        candidates with an unstable std::sort -- the greedy assignment then depends on the sort's tie order (the
        restatement freezes score desc, idx1, idx2; seen on seeds 114 / 117 with noise 0: peaks identical, two
        limbs assigned differently among equal-score candidates);
      * feature maps wider than 4:1 (the default resolution then shrinks an axis) are covered separately by
        test_oracle_matches_reference_golden_on_shrinking_resolutions."""
    rng = np.random.default_rng(seed)
    hf = int(rng.integers(9, 50))
    wf = int(rng.integers(9, min(90, 4 * hf) + 1))
    P = int(rng.integers(0, 13))
    noise = float(rng.choice([0.005, 0.02, 0.06]))
    conf, paf = syn.make_frame_tensors(seed, P, hf, wf, noise=noise)
    ct, pt = float(rng.choice([0.05, 0.1, 0.3])), float(rng.choice([0.02, 0.05, 0.2]))
    if seed % 3 == 0:      # user resolution (w, h) >= the feature map
        rw, rh = int(hf * rng.uniform(1.0, 4.6)) + 1, int(wf * rng.uniform(1.0, 4.6)) + 1
        rw, rh = max(rw, wf), max(rh, hf)
    else:
        rw = rh = -1
    rp = oracle.RefParser(ct, pt, rw, rh)
    want = rp.process(conf, paf)
    rp.close()
    got = oracle.oracle_process(conf, paf, ct, pt, rw, rh)["humans"]
    assert got.tobytes() == want.tobytes(), (hf, wf, P, noise, ct, pt, rw, rh, len(got), len(want))
