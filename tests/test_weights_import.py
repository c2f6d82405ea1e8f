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


# ---------------------------------------------------------------------------------------------------------------------
# MobilenetThin-OpenPose (hyperpose/Model/backbones.py:233-297, openpose/model/mbv2_th_openpose.py:36-177)
# ---------------------------------------------------------------------------------------------------------------------
def _tl_arrays(order, seed):
    """random arrays in TensorLayer's all_weights order / layouts for a *_layer_order list"""
    rng = np.random.default_rng(seed)
    out = []
    for kind, name, co, ci, k in order:
        if kind in ("conv", "conv_nobias"):
            out.append((rng.standard_normal((k, k, ci, co)) * np.sqrt(2.0 / (ci * k * k))).astype(np.float32))     # HWIO
            if kind == "conv":
                out.append((rng.standard_normal(co) * 0.05).astype(np.float32))
        elif kind == "dwconv":
            out.append((rng.standard_normal((k, k, co, 1)) * np.sqrt(2.0 / (k * k))).astype(np.float32))          # [kh, kw, C, 1]
        elif kind == "bn":                                                                                         # beta, gamma, moving_mean, moving_var
            out += [rng.normal(0, 0.1, co).astype(np.float32), rng.uniform(0.7, 1.3, co).astype(np.float32),
                    rng.normal(0, 0.1, co).astype(np.float32), rng.uniform(0.6, 1.4, co).astype(np.float32)]
        elif kind == "prelu":
            out.append((rng.standard_normal(co) * 0.5).astype(np.float32))
    return out


def _same_pad(x, k, stride):
    """TensorFlow 'SAME': out = ceil(in / stride), the odd pixel of padding goes AFTER"""
    pads = []
    for n in (x.shape[3], x.shape[2]):
        total = max((-(-n // stride) - 1) * stride + k - n, 0)
        pads += [total // 2, total - total // 2]
    return F.pad(x, pads)


class _TlReader:
    """consumes an all_weights list the way the reference's layers would"""

    def __init__(self, arrays):
        self.it = iter(arrays)

    def conv(self, x, stride=1, bias=True):
        f = next(self.it)
        b = torch.from_numpy(next(self.it)) if bias else None
        return F.conv2d(_same_pad(x, f.shape[0], stride), torch.from_numpy(f.transpose(3, 2, 0, 1).copy()), b, stride=stride)

    def dwconv(self, x, stride=1):
        f = next(self.it)                                       # [kh, kw, C, 1]
        w = torch.from_numpy(f.transpose(2, 3, 0, 1).copy())    # [C, 1, kh, kw]
        return F.conv2d(_same_pad(x, f.shape[0], stride), w, None, stride=stride, groups=f.shape[2])

    def bn(self, x):
        beta, gamma, mean, var = (torch.from_numpy(next(self.it)).view(1, -1, 1, 1) for _ in range(4))
        return (x - mean) / torch.sqrt(var + 1e-5) * gamma + beta

    def prelu(self, x):
        return F.prelu(x, torch.sigmoid(torch.from_numpy(next(self.it))))

    def separable(self, x, stride=1, act=True):
        a = F.relu if act else (lambda t: t)
        x = a(self.bn(self.dwconv(x, stride)))
        return a(self.bn(self.conv(x, bias=False)))

    def done(self):
        assert next(self.it, None) is None


def _mobilenet_thin_reference(arrays, x, n_stages):
    r = _TlReader(arrays)
    x = F.relu(r.bn(F.relu(r.conv(x, stride=2))))                                          # conv_block: Conv2d(act=relu) + BatchNorm(act=relu)
    strides = [1, 2, 1, 2, 1, 1, 1, 1, 1, 1, 1]                                            # scale_size 8 (backbones.py:258-275)
    cat = []
    for i, st in enumerate(strides, start=1):
        x = r.separable(x, st)
        if i == 3:
            cat.append(F.max_pool2d(x, 2, 2, ceil_mode=True))
        if i in (7, 11):
            cat.append(x)
    feat = torch.cat(cat, 1)

    def branch(t):
        for k in range(5):
            t = r.separable(t, 1, act=(k < 4))                                             # the last block is built with act=None
        return t

    conf, paf = branch(feat), branch(feat)
    for _ in range(1, n_stages):
        t = torch.cat([feat, conf, paf], 1)
        conf, paf = branch(t), branch(t)
    r.done()
    return conf, paf


@pytest.mark.parametrize("n_stages", [1, 3])
def test_mobilenet_thin_import_equals_reference_definition(n_stages):
    arrays = _tl_arrays(W.mobilenet_thin_layer_order(n_stages), 21)
    g = models.mobilenet_thin_openpose(n_stages=n_stages, weights=W.MobilenetThinWeights(arrays, n_stages))
    frames = np.random.default_rng(4).integers(0, 256, (2, 64, 88, 3), dtype=np.uint8)
    conf, paf, _ = torch_ref.run_graph(g, frames, flip_rgb=True, device="cpu")
    x = torch.from_numpy(np.ascontiguousarray((frames.astype(np.float64) / 255).astype(np.float32)[..., ::-1].transpose(0, 3, 1, 2)))
    rc, rp = _mobilenet_thin_reference(arrays, x, n_stages)
    assert conf.shape == rc.shape == (2, 19, 8, 11) and paf.shape == rp.shape == (2, 38, 8, 11)
    tol = 2e-4 * max(1.0, float(rc.abs().max()), float(rp.abs().max()))
    assert float((conf.cpu() - rc).abs().max()) < tol and float((paf.cpu() - rp).abs().max()) < tol


def test_bn_net_weights_reject_wrong_lists():
    order = W.mobilenet_thin_layer_order(1)
    arrays = _tl_arrays(order, 2)
    with pytest.raises(ValueError):
        W.MobilenetThinWeights(arrays[:-1], 1)
    with pytest.raises(ValueError):
        W.MobilenetThinWeights(arrays + [np.zeros(3, np.float32)], 1)
    bad = list(arrays); bad[0] = bad[0].transpose(3, 2, 0, 1)      # OIHW instead of HWIO
    with pytest.raises(ValueError):
        W.MobilenetThinWeights(bad, 1)


# ---------------------------------------------------------------------------------------------------------------------
# ResNet-50 networks: LightWeightOpenPose head (backbones.py:587-698 + lw_openpose.py:33-191) and PifPaf (pifpaf/model.py:41-281)
# ---------------------------------------------------------------------------------------------------------------------
def _resnet50_reference(r, x, layout, use_pool, eps):
    def bn(t):
        beta, gamma, mean, var = (torch.from_numpy(next(r.it)).view(1, -1, 1, 1) for _ in range(4))
        return (t - mean) / torch.sqrt(var + eps) * gamma + beta

    x = F.relu(bn(r.conv(x, stride=2, bias=False)))                                         # conv1 7x7/2 + bn1
    if use_pool:                                                                            # MaxPool2d(3, 2, 'SAME'): padding never wins
        n_h, n_w = x.shape[2], x.shape[3]
        pads = []
        for n in (n_w, n_h):
            total = max((-(-n // 2) - 1) * 2 + 3 - n, 0)
            pads += [total // 2, total - total // 2]
        x = F.max_pool2d(F.pad(x, pads, value=float("-inf")), 3, 2)
    cin = 64
    for nf, nblk, st0 in layout:
        for k in range(1, nblk + 1):
            st = st0 if k == 1 else 1
            res = x
            if st != 1 or cin != 4 * nf:                                                    # `downsample` is created (and saved) before main_block
                res = bn(r.conv(x, stride=st, bias=False))
            y = F.relu(bn(r.conv(x, bias=False)))
            y = F.relu(bn(r.conv(y, stride=st, bias=False)))
            y = bn(r.conv(y, bias=False))
            x = F.relu(y + res)
            cin = 4 * nf
    return x


def test_resnet50_lw_openpose_import_equals_reference_definition():
    arrays = _tl_arrays(W.resnet50_lw_layer_order(), 31)
    g = models.resnet50_lw_openpose(weights=W.Resnet50LwWeights(arrays))
    frames = np.random.default_rng(5).integers(0, 256, (1, 64, 80, 3), dtype=np.uint8)
    conf, paf, _ = torch_ref.run_graph(g, frames, flip_rgb=True, device="cpu")
    x = torch.from_numpy(np.ascontiguousarray((frames.astype(np.float64) / 255).astype(np.float32)[..., ::-1].transpose(0, 3, 1, 2)))
    r = _TlReader(arrays)
    feat = _resnet50_reference(r, x, [(64, 3, 1), (128, 4, 2), (256, 6, 1), (512, 3, 1)], True, 1e-5)
    cb = lambda t: F.relu(r.bn(r.conv(t)))                                                  # conv_block: Conv2d(+bias), BatchNorm(relu)
    t = F.relu(r.conv(feat))                                                                # Cpm_stage
    t = t + cb(cb(cb(t)))
    cpm = F.relu(r.conv(t))
    t = F.relu(r.conv(F.relu(r.conv(F.relu(r.conv(cpm))))))                                 # Init_stage.main_block
    conf_r = r.conv(F.relu(r.conv(t)))
    paf_r = r.conv(F.relu(r.conv(t)))
    t = torch.cat([cpm, conf_r, paf_r], 1)
    for _ in range(5):                                                                      # Refinement_block x 5
        t = F.relu(r.conv(t))
        t = t + cb(cb(t))
    conf_r = r.conv(F.relu(r.conv(t)))
    paf_r = r.conv(F.relu(r.conv(t)))
    r.done()
    assert conf.shape == conf_r.shape == (1, 19, 8, 10) and paf.shape == paf_r.shape == (1, 38, 8, 10)
    tol = 2e-4 * max(1.0, float(conf_r.abs().max()), float(paf_r.abs().max()))
    assert float((conf.cpu() - conf_r).abs().max()) < tol and float((paf.cpu() - paf_r).abs().max()) < tol


def test_resnet50_pifpaf_import_equals_reference_definition():
    arrays = _tl_arrays(W.resnet50_pifpaf_layer_order(), 41)
    g = models.resnet50_pifpaf(weights=W.Resnet50PifPafWeights(arrays))
    frames = np.random.default_rng(6).integers(0, 256, (1, 64, 96, 3), dtype=np.uint8)
    _, _, bufs = torch_ref.run_graph(g, frames, flip_rgb=True, device="cpu")
    head = next(op for op in g.ops if op.type == models.OP_PIFPAF_HEAD)
    pif_raw, paf_raw = bufs[head.in_buf].cpu(), bufs[head.res_buf].cpu()                     # the 1x1 heads before pixel-shuffle / activations
    x = torch.from_numpy(np.ascontiguousarray((frames.astype(np.float64) / 255).astype(np.float32)[..., ::-1].transpose(0, 3, 1, 2)))
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1); std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    r = _TlReader(arrays)
    feat = _resnet50_reference(r, (x - mean) / std, [(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)], False, 1e-4)
    pif_r, paf_r = r.conv(feat), r.conv(feat)
    r.done()
    assert pif_r.shape == (1, 340, 4, 6) and paf_r.shape == (1, 684, 4, 6)
    tol = 3e-4 * max(1.0, float(pif_r.abs().max()), float(paf_r.abs().max()))
    assert float((pif_raw[:, :340] - pif_r).abs().max()) < tol and float((paf_raw[:, :684] - paf_r).abs().max()) < tol
