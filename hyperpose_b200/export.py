"""Writes an HPB2PACK model pack (the file `tensorrt(tensorrt_serialized{path}, ...)` / `hp_engine_create` load):

    python -m hyperpose_b200.export --model openpose_vgg19 --out vgg19.pack [--weights trained.npz] [--seed 0]

`--weights` is a TensorLayer `save_weights(format="npz")` file of the reference's model of that architecture -- OpenPose-VGG19 (also
the name-keyed `npz_dict` form), MobilenetThin-OpenPose, LightWeightOpenPose on ResNet-50, PifPaf on ResNet-50; hyperpose_b200/weights.py
spells out the all_weights order of each, BatchNorm statistics are folded.  Without it the pack holds seeded random weights, which is
what the benchmarks and tests use (no trained model can be downloaded offline).  Replaces the .onnx / .uff / .trt files of
include/hyperpose/utility/model.hpp:13-32."""
from __future__ import annotations

import argparse

from . import models, weights


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--model", default="openpose_vgg19",
                    choices=["openpose_vgg19", "mobilenet_thin_openpose", "resnet50_lw_openpose", "resnet50_pifpaf", "tiny_test_net"])
    ap.add_argument("--out", required=True)
    ap.add_argument("--weights", default=None, help="TensorLayer save_weights(format='npz') file of the same architecture")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args(argv)
    if a.weights:
        loaders = {"openpose_vgg19": weights.ListWeights, "mobilenet_thin_openpose": weights.MobilenetThinWeights,
                   "resnet50_lw_openpose": weights.Resnet50LwWeights, "resnet50_pifpaf": weights.Resnet50PifPafWeights}
        if a.model not in loaders:
            ap.error(f"--weights: no trained-weight layout for {a.model}")
        g = getattr(models, a.model)(weights=loaders[a.model].from_npz(a.weights))
    else:
        g = getattr(models, a.model)(a.seed)
    blob = g.to_pack()
    with open(a.out, "wb") as f:
        f.write(blob)
    print(f"{a.out}: {g.name}, {len(g.ops)} ops, {len(blob) / 1e6:.1f} MB")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
