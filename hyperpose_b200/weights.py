"""Weights of the reference's OpenPose-VGG19 model (hyperpose/Model/openpose/model/openpose.py + backbones.py:447-509)
for hyperpose_b200.models.openpose_vgg19 -- SURVEY.md 8f rank 1, the exporter side of the model pack.

The reference saves a trained model with TensorLayer's `Model.save_weights(path, format="npz")`: an ORDERED list of
arrays = `model.all_weights`, i.e. layer-creation order; most layers of the stages carry auto-generated names
(openpose.py:36-39,126-149), so position -- not name -- is the stable key, exactly how `tl.files.load_and_assign_npz`
restores them.  `openpose_vgg19_layer_order()` spells that order out with the array shapes (TensorFlow layouts:
Conv2d filters HWIO, biases [O], PRelu alpha with `in_channels` elements) and `ListWeights` consumes such a list,
checking every shape, and hands the layers out by the names the graph builder uses.

Layer names: VGG `conv1_1 .. conv4_2`, `cpm_1/2`, `init.{conf|paf}.{1..5}` and `ref{1..5}.{conf|paf}.{1..7}`.
"""
from __future__ import annotations

import numpy as np


def openpose_vgg19_layer_order(n_stages: int = 6, n_conf: int = 19, n_paf: int = 38):
    """[(kind, name, cout, cin, k)] in `all_weights` order; kind 'conv' contributes (filters, biases), 'prelu' (alpha,)."""
    order = []
    vgg = [("conv1_1", 64, 3), ("conv1_2", 64, 64), ("conv2_1", 128, 64), ("conv2_2", 128, 128), ("conv3_1", 256, 128),
           ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv3_4", 256, 256), ("conv4_1", 512, 256), ("conv4_2", 512, 512)]
    for name, co, ci in vgg:                                     # backbones.py:461-476 (ReLU inside the Conv2d: no extra weight)
        order.append(("conv", name, co, ci, 3))
    order += [("conv", "cpm_1", 256, 512, 3), ("conv", "cpm_2", 128, 256, 3)]   # openpose.py:36-39

    def stage(prefix, cin, layers):
        for branch, cout_last in (("conf", n_conf), ("paf", n_paf)):            # conf_block is created before paf_block (:126,:138)
            c_prev = cin
            for i, (co, k) in enumerate(layers, start=1):
                co = cout_last if co is None else co
                order.append(("conv", f"{prefix}.{branch}.{i}", co, c_prev, k))
                order.append(("prelu", f"{prefix}.{branch}.{i}", co, 0, 0))     # tl.layers.PRelu after every conv
                c_prev = co

    stage("init", 128, [(128, 3), (128, 3), (128, 3), (512, 1), (None, 1)])     # openpose.py:119-154
    for s in range(1, n_stages):                                                # openpose.py:156-199
        stage(f"ref{s}", 128 + n_conf + n_paf, [(128, 7)] * 5 + [(128, 1), (None, 1)])
    return order


class RandomWeights:
    """He-normal convs, small biases, PReLU slopes in [0.1, 0.4) -- what bench.py and the tests run on (no trained
    weights can be fetched offline).  Deterministic per (seed, layer name)."""

    def __init__(self, seed: int = 0):
        self.seed = seed

    def _rng(self, name: str, salt: int):
        return np.random.default_rng([self.seed, salt] + [ord(c) for c in name])

    def conv(self, name, cout, cin, k, gain=2.0):
        r = self._rng(name, 1)
        w = (r.standard_normal((cout, cin, k, k)) * np.sqrt(gain / (cin * k * k))).astype(np.float32)
        return w, (r.standard_normal(cout) * 0.05).astype(np.float32)

    def prelu(self, name, c):
        return self._rng(name, 2).uniform(0.1, 0.4, c).astype(np.float32)


class ListWeights:
    """An `all_weights`-ordered list of arrays (TensorFlow layouts) for openpose_vgg19."""

    def __init__(self, arrays, n_stages: int = 6):
        arrays = [np.asarray(a) for a in arrays]
        self._conv, self._prelu = {}, {}
        it = iter(arrays)
        for kind, name, co, ci, k in openpose_vgg19_layer_order(n_stages):
            try:
                if kind == "conv":
                    f, b = next(it), next(it)
                    if f.shape != (k, k, ci, co) or b.size != co:
                        raise ValueError(f"{name}: expected filters {(k, k, ci, co)} + {co} biases, got {f.shape} / {b.shape}")
                    self._conv[name] = (np.ascontiguousarray(f.transpose(3, 2, 0, 1)).astype(np.float32), b.reshape(co).astype(np.float32))
                else:
                    a = next(it)
                    if a.size != co:
                        raise ValueError(f"{name}: expected {co} PRelu slopes, got shape {a.shape}")
                    # TensorLayer 2.2.3 (requirements.txt:6) constrains the slope: PRelu.build creates
                    # `alpha_var_constrained = tf.nn.sigmoid(alpha_var)` and forward computes relu(x) - sigmoid(alpha) * relu(-x);
                    # the SAVED array is the raw variable (init ~N(0, 0.05) => slope ~0.5), so the slope the pack needs is sigmoid(alpha).
                    self._prelu[name] = (1.0 / (1.0 + np.exp(-a.reshape(co).astype(np.float64)))).astype(np.float32)
            except StopIteration:
                raise ValueError(f"weight list ends before {kind} {name}") from None
        if next(it, None) is not None:
            raise ValueError("weight list is longer than the model")

    @classmethod
    def from_npz(cls, path: str, n_stages: int = 6):
        """Either TensorLayer weight file of the reference:
          * `save_weights(format="npz")` / `tl.files.save_npz`: one object array under the key 'params', in all_weights order
            (hyperpose/Model/train.py:582);
          * `save_weights(format="npz_dict")` / `tl.files.save_npz_dict`: one entry per weight, keyed by the weight's name
            `<layer name>/<filters|biases|alpha>:0` (train.py:319, eval.py:109) -> `from_name_dict`."""
        # npz_dict files need no pickle; only the legacy `params` object array does.  Unpickling executes code from the file:
        # it is re-opened with allow_pickle=True only for that key, and only a file you trust should be passed here.
        z = np.load(path, allow_pickle=False)
        if "params" in z.files:
            z = np.load(path, allow_pickle=True)
            return cls(list(z["params"]), n_stages)
        return cls.from_name_dict({k: z[k] for k in z.files}, n_stages)

    @classmethod
    def from_name_dict(cls, named: dict, n_stages: int = 6):
        """Name-keyed weights -> the all_weights order.  The VGG layers carry explicit names (`conv1_1` .. `conv4_2`,
        backbones.py:461-476); every other layer of the model is created without a name (openpose.py:36-39,126-149) and gets
        TensorLayer's automatic `<class>_<counter>` (`conv2d_7`, `prelu_3`), the counter running in creation order per layer
        class -- so within a class, ascending counter == creation order, which is all that is needed to line the entries up
        with `openpose_vgg19_layer_order`.  Shapes are checked entry by entry by the constructor."""
        import re
        layers = {}
        for key, arr in named.items():
            # only the three weight kinds of the model; anything else in the file (optimizer slots, counters) is ignored
            if not key.endswith(("/filters:0", "/biases:0", "/alpha:0")):
                continue
            lname = key.split("/")[0]
            layers.setdefault(lname, []).append(np.asarray(arr))

        def split(lname):      # (filters, biases) of a conv layer | (alpha,) of a PRelu
            arrs = layers[lname]
            four = [a for a in arrs if a.ndim == 4]
            one = [a for a in arrs if a.ndim != 4]
            if len(four) == 1 and len(one) == 1:
                return [four[0], one[0]]
            if not four and len(one) == 1:
                return [one[0]]
            raise ValueError(f"layer {lname}: cannot tell filters / biases / alpha apart ({[a.shape for a in arrs]})")

        def counter(lname):
            m = re.search(r"_(\d+)$", lname)
            return int(m.group(1)) if m else 0

        order = openpose_vgg19_layer_order(n_stages)
        vgg_names = [name for kind, name, *_ in order if kind == "conv" and name.startswith("conv") and "." not in name]
        unnamed_conv = sorted((n for n in layers if n not in vgg_names and any(a.ndim == 4 for a in layers[n])), key=counter)
        unnamed_prelu = sorted((n for n in layers if n not in vgg_names and all(a.ndim != 4 for a in layers[n])), key=counter)
        ci, pi = iter(unnamed_conv), iter(unnamed_prelu)
        out = []
        try:
            for kind, name, *_ in order:
                if kind == "conv":
                    out += split(name if name in vgg_names else next(ci))
                else:
                    out += split(next(pi))
        except (StopIteration, KeyError) as e:
            raise ValueError(f"name-keyed weight file does not hold the layers of OpenPose-VGG19 with {n_stages} stages ({e!r})") from None
        if next(ci, None) is not None or next(pi, None) is not None:
            raise ValueError("name-keyed weight file holds more conv / PRelu layers than the model")
        return cls(out, n_stages)

    def conv(self, name, cout, cin, k, gain=2.0):
        w, b = self._conv[name]
        assert w.shape == (cout, cin, k, k), (name, w.shape)
        return w, b

    def prelu(self, name, c):
        return self._prelu[name]
