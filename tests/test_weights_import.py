"""Weight import for OpenPose-VGG19 and an INDEPENDENT check of the merged-branch graph:
a TensorLayer-ordered weight list (HWIO filters, layer-creation order of hyperpose/Model/openpose/model/openpose.py)
is loaded by hyperpose_b200.weights.ListWeights into models.openpose_vgg19, and the graph -- which runs the conf and paf
branches of every stage as ONE merged / grouped / block-diagonal conv -- must reproduce a plain PyTorch model that is
written straight from the reference definition: two separate branches per stage and an explicit concat."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from hyperpose_b200 import models, weights as W
from oracle import torch_backbone as torch_ref


def _tl_weight_list(seed, n_stages):
    """random weights in TensorLayer's all_weights order and layouts"""
    rng = np.random.default_rng(seed)
    out = []
    for kind, name, co, ci, k in W.openpose_vgg19_layer_order(n_stages):
        if kind == "conv":
            out.append((rng.standard_normal((k, k, ci, co)) * np.sqrt(2.0 / (ci * k * k))).astype(np.float32))   # HWIO
            out.append((rng.standard_normal(co) * 0.05).astype(np.float32))
        else:
            out.append((rng.standard_normal(co) * 0.5).astype(np.float32))   # raw alpha variable (any sign)
    return out


def _reference_forward(arrays, x, n_stages):
    """openpose.py:57-88 + backbones.py:447-509 in plain torch, consuming `arrays` in all_weights order; x: [N,3,H,W]"""
    it = iter(arrays)

    def conv(x, act=None):
        f, b = next(it), next(it)
        k = f.shape[0]
        y = F.conv2d(x, torch.from_numpy(f.transpose(3, 2, 0, 1).copy()), torch.from_numpy(b), padding=k // 2)
        return F.relu(y) if act == "relu" else y

    def prelu(x):
        # TensorLayer 2.2.3 PRelu: relu(x) - sigmoid(alpha) * relu(-x) on the RAW saved alpha
        return F.prelu(x, torch.sigmoid(torch.from_numpy(next(it))))

    def pool(x):   # MaxPool2d(2, 2), TF 'SAME': window clipped at the border
        return F.max_pool2d(x, 2, 2, ceil_mode=True)

    x = x - torch.tensor([103.939, 116.779, 123.68]).view(1, 3, 1, 1) / 255
    x = conv(conv(x, "relu"), "relu"); x = pool(x)
    x = conv(conv(x, "relu"), "relu"); x = pool(x)
    x = conv(conv(conv(conv(x, "relu"), "relu"), "relu"), "relu"); x = pool(x)
    x = conv(conv(x, "relu"), "relu")
    feat = conv(conv(x, "relu"), "relu")                       # cpm_stage

    def branch(x, n):
        for _ in range(n):
            x = prelu(conv(x))
        return x

    conf, paf = branch(feat, 5), branch(feat, 5)               # conf_block first, then paf_block
    for _ in range(1, n_stages):
        rx = torch.cat([feat, conf, paf], 1)
        conf, paf = branch(rx, 7), branch(rx, 7)
    assert next(it, None) is None
    return conf, paf


@pytest.mark.parametrize("n_stages", [2, 6])
def test_merged_graph_equals_two_branch_reference(n_stages):
    torch.manual_seed(0)
    arrays = _tl_weight_list(3, n_stages)
    g = models.openpose_vgg19(n_stages=n_stages, weights=W.ListWeights(arrays, n_stages))
    frames = np.random.default_rng(1).integers(0, 256, (2, 40, 56, 3), dtype=np.uint8)
    conf, paf, _ = torch_ref.run_graph(g, frames, flip_rgb=True, device="cpu")
    conf, paf = conf.cpu(), paf.cpu()
    x = torch.from_numpy(np.ascontiguousarray((frames.astype(np.float64) / 255).astype(np.float32)[..., ::-1].transpose(0, 3, 1, 2)))
    rc, rp = _reference_forward(arrays, x, n_stages)
    assert conf.shape == rc.shape == (2, 19, 5, 7) and paf.shape == rp.shape == (2, 38, 5, 7)
    tol = 1e-4 * max(1.0, float(rc.abs().max()), float(rp.abs().max()))
    assert float((conf - rc).abs().max()) < tol and float((paf - rp).abs().max()) < tol


def test_list_weights_rejects_wrong_shapes_and_lengths():
    arrays = _tl_weight_list(0, 2)
    with pytest.raises(ValueError):
        W.ListWeights(arrays[:-1], 2)
    with pytest.raises(ValueError):
        W.ListWeights(arrays + [np.zeros(3, np.float32)], 2)
    bad = list(arrays); bad[0] = bad[0].transpose(3, 2, 0, 1)      # OIHW instead of HWIO
    with pytest.raises(ValueError):
        W.ListWeights(bad, 2)


def test_npz_round_trip(tmp_path):
    arrays = _tl_weight_list(5, 2)
    obj = np.empty(len(arrays), dtype=object)
    for i, a in enumerate(arrays):
        obj[i] = a
    np.savez(tmp_path / "model.npz", params=obj)                   # tl.files.save_npz layout
    ws = W.ListWeights.from_npz(str(tmp_path / "model.npz"), 2)
    g1 = models.openpose_vgg19(n_stages=2, weights=ws)
    g2 = models.openpose_vgg19(n_stages=2, weights=W.ListWeights(arrays, 2))
    assert g1.to_pack() == g2.to_pack()


def test_random_weights_are_deterministic():
    assert models.openpose_vgg19(seed=7, n_stages=2).to_pack() == models.openpose_vgg19(seed=7, n_stages=2).to_pack()
    assert models.openpose_vgg19(seed=7, n_stages=2).to_pack() != models.openpose_vgg19(seed=8, n_stages=2).to_pack()


@pytest.mark.gpu
def test_engine_on_imported_weights_matches_two_branch_reference():
    """the CUDA engine (fp16 operands, fp32 accumulation) on an imported weight list vs the two-branch fp32 torch model"""
    from hyperpose_b200 import capi
    n_stages = 6
    arrays = _tl_weight_list(11, n_stages)
    g = models.openpose_vgg19(n_stages=n_stages, weights=W.ListWeights(arrays, n_stages))
    H, Wd, N = 64, 96, 2
    frames = np.random.default_rng(2).integers(0, 256, (N, H, Wd, 3), dtype=np.uint8)
    eng = capi.Engine(g.to_pack(), (Wd, H), max_batch_size=N)
    eng.infer_u8(frames)
    conf, paf = eng.read_outputs(N)
    eng.close()
    x = torch.from_numpy(np.ascontiguousarray((frames.astype(np.float64) / 255).astype(np.float32)[..., ::-1].transpose(0, 3, 1, 2)))
    rc, rp = _reference_forward(arrays, x, n_stages)
    for got, want in ((conf, rc.numpy()), (paf, rp.numpy())):
        assert got.shape == want.shape
        assert float(np.abs(got - want).max()) <= 3e-2 * max(1.0, float(np.abs(want).max()))


def test_name_keyed_weight_file_gives_the_same_model(tmp_path):
    """`save_weights(format="npz_dict")` (hyperpose/Model/train.py:319): entries keyed `<layer>/<filters|biases|alpha>:0`, the
    stage layers under TensorLayer's automatic `conv2d_<n>` / `prelu_<n>` names; key order in the file is arbitrary"""
    n_stages = 3
    arrays = _tl_weight_list(5, n_stages)
    it = iter(arrays)
    named, nc, npr = {}, 0, 0
    for kind, name, co, ci, k in W.openpose_vgg19_layer_order(n_stages):
        if kind == "conv":
            if "." in name or name.startswith("cpm"):
                nc += 1
                lname = f"conv2d_{nc + 7}"        # counters do not start at 1: other models were built earlier in the process
            else:
                lname = name
            named[f"{lname}/filters:0"] = next(it)
            named[f"{lname}/biases:0"] = next(it)
        else:
            npr += 1
            named[f"prelu_{npr + 2}/alpha:0"] = next(it)
    keys = list(named)
    np.random.default_rng(0).shuffle(keys)
    path = tmp_path / "newest_model.npz"
    np.savez(path, **{k: named[k] for k in keys})
    a = W.ListWeights.from_npz(str(path), n_stages)
    b = W.ListWeights(arrays, n_stages)
    for kind, name, co, ci, k in W.openpose_vgg19_layer_order(n_stages):
        if kind == "conv":
            assert np.array_equal(a.conv(name, co, ci, k)[0], b.conv(name, co, ci, k)[0]) and np.array_equal(a.conv(name, co, ci, k)[1], b.conv(name, co, ci, k)[1])
        else:
            assert np.array_equal(a.prelu(name, co), b.prelu(name, co))
    # a file with a layer missing is rejected, not silently shifted
    del named["prelu_4/alpha:0"]
    with pytest.raises(ValueError):
        W.ListWeights.from_name_dict(named, n_stages)
