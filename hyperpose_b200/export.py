"""Writes an HPB2PACK model pack (the file `tensorrt(tensorrt_serialized{path}, ...)` / `hp_engine_create` load):

    python -m hyperpose_b200.export --model openpose_vgg19 --out vgg19.pack [--weights trained.npz] [--seed 0]

`--weights` is a TensorLayer `save_weights(format="npz")` file of the reference's OpenPose-VGG19 model
(hyperpose_b200/weights.py explains the order); without it the pack holds seeded random weights, which is what the
benchmarks and tests use (no trained model can be downloaded offline).  Replaces the .onnx / .uff / .trt files of
include/hyperpose/utility/model.hpp:13-32."""
from __future__ import annotations

import argparse

from . import models, weights


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--model", default="openpose_vgg19",
                    choices=["openpose_vgg19", "mobilenet_thin_openpose", "resnet50_lw_openpose", "resnet50_pifpaf", "tiny_test_net"])
    ap.add_argument("--out", required=True)
    ap.add_argument("--weights", default=None, help="TensorLayer npz weight list (openpose_vgg19 only)")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args(argv)
    if a.weights:
        if a.model != "openpose_vgg19":
            ap.error("--weights is implemented for openpose_vgg19")
        g = models.openpose_vgg19(weights=weights.ListWeights.from_npz(a.weights))
    else:
        g = getattr(models, a.model)(a.seed)
    blob = g.to_pack()
    with open(a.out, "wb") as f:
        f.write(blob)
    print(f"{a.out}: {g.name}, {len(g.ops)} ops, {len(blob) / 1e6:.1f} MB")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
