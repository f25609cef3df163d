"""Drop-in for the reference's generate_gp_training_data_cifar.py (hot loop `eval_superpixel()` :236-342).

    python generate_gp_training_data_cifar.py --arch resnet --depth 56 [--num_mask_samples 1000]

Loads ./saved_checkpoints/cifar10+-resnet-56/model_best.pth.tar through `models.<arch>.createModel` +
nn.DataParallel exactly like :71-79,:248-250, then removes 5 random superpixels per mask, min-max renormalises,
scores (:307-333) — all masks at once on the B200.  The CIFAR-10 test set is not available offline: the image is a
seeded synthetic 32x32 unless --image is given."""
from __future__ import annotations

import argparse
import os
from importlib import import_module

import numpy as np
import torch
import torch.nn as nn

from network_interpretation_imagenet_b200 import synthetic
from network_interpretation_imagenet_b200.pipeline import run_generator

parser = argparse.ArgumentParser(description="CIFAR perturbation-data generator (B200 engine)")
parser.add_argument("--arch", "-a", default="resnet", choices=["resnet", "densenet"])
parser.add_argument("--depth", "-d", default=56, type=int)
parser.add_argument("--data", default="cifar10", choices=["cifar10", "cifar100"])
parser.add_argument("--death-mode", default="none", choices=["none", "linear", "uniform"])
parser.add_argument("--death-rate", default=0.5, type=float)
parser.add_argument("--growth-rate", "-g", default=12, type=int)
parser.add_argument("--bn-size", default=4, type=int)
parser.add_argument("--compression", default=0.5, type=float)
parser.add_argument("--drop-rate", default=0.0, type=float)
parser.add_argument("--resume", default="./saved_checkpoints/cifar10+-resnet-56/model_best.pth.tar")
parser.add_argument("--num_mask_samples", default=1000, type=int)
parser.add_argument("--image", default=None, help=".npy file holding a 3x32x32 fp32 image in loader space")
parser.add_argument("--target", default=None, type=int)
parser.add_argument("--mask-seed", default=0, type=int)
parser.add_argument("--precision", default="x3", choices=["bf16", "fp32", "x3"],
                    help="fp32: CUDA-core parity mode (1e-4); x3: split-bf16 products on the tensor cores, fp32-grade (1e-4); "
                         "bf16: tcgen05 path (1e-2) with the device-side tie policy")
parser.add_argument("--batch-size", "-b", default=512, type=int)
parser.add_argument("--no-write", action="store_true")


def getModel(arch, **kargs):
    """:71-79 — import_module('models.'+arch).createModel(**kargs) wrapped in DataParallel."""
    m = import_module("models." + arch)
    model = m.createModel(**kargs)
    return nn.DataParallel(model).cuda()


def eval_superpixel(args):
    num_classes = 10 if args.data == "cifar10" else 100
    model = getModel(args.arch, depth=args.depth, data=args.data, num_classes=num_classes, death_mode=args.death_mode,
                     death_rate=args.death_rate, growth_rate=args.growth_rate, bn_size=args.bn_size,
                     compression=args.compression, drop_rate=args.drop_rate)
    if args.resume and os.path.isfile(args.resume):
        checkpoint = torch.load(args.resume, map_location="cpu", weights_only=False)
        model.load_state_dict(checkpoint["state_dict"])            # :249-250
        print("=> loaded checkpoint '{}' (epoch {})".format(args.resume, checkpoint.get("epoch")))
    else:
        print("=> no checkpoint found at '{}': random init".format(args.resume))
    model.eval()
    model.module.configure_engine(precision=args.precision, max_batch=args.batch_size)
    image = np.load(args.image).astype(np.float32) if args.image else synthetic.synthetic_image("cifar")
    output = model.module(torch.from_numpy(image)[None].cuda())     # :301-302
    pred = int(output.argmax(1)[0])
    target = pred if args.target is None else args.target
    # bf16: hand over the torch module so the engine can also lower the fp32 copy its tie policy re-scores with
    net = model.module if args.precision == "bf16" else model.module.engine((32, 32))
    res = run_generator("cifar", net, image, target, args.num_mask_samples, args.mask_seed,
                        precision=args.precision, max_batch=args.batch_size,
                        mask_dir=None if args.no_write else "./masks")
    return res["correct_pred_count"], res["wrong_pred_count"]


if __name__ == "__main__":
    eval_superpixel(parser.parse_args())
