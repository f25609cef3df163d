"""Drop-in for the reference's generate_gp_training_data_mnist.py (hot loop `eval_superpixel()` :153-269).

    python generate_gp_training_data_mnist.py [--num_mask_samples 1000]

`Classification_Net` keeps the reference's parameter names (so ./saved_checkpoints/mnist/checkpoint.pth.tar loads
under key 'model', :157-158) and returns the 4-tuple (x0, x1, x2, pred0) (:97-105); its forward runs in libnib.so.
Per mask: one random superpixel removed (plus the reference's unused randint draw, :211), min-max renormalise,
x 1/255, forward, softmax max-prob / arg-max (:203-259).  MNIST cannot be downloaded offline: the image is a seeded
synthetic 28x28 unless --image is given."""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch
import torch.nn as nn

from models._engine_module import EngineModule
from network_interpretation_imagenet_b200 import synthetic
from network_interpretation_imagenet_b200.pipeline import run_generator

parser = argparse.ArgumentParser(description="MNIST perturbation-data generator (B200 engine)")
parser.add_argument("--batch-size", type=int, default=64)
parser.add_argument("--test-batch-size", type=int, default=1)
parser.add_argument("--seed", type=int, default=1)
parser.add_argument("--num_mask_samples", type=int, default=1000)
parser.add_argument("--image", default=None, help=".npy file holding a 1x28x28 fp32 image in [0,1]")
parser.add_argument("--target", default=None, type=int)
parser.add_argument("--mask-seed", default=0, type=int)
parser.add_argument("--precision", default="fp32", choices=["bf16", "fp32", "x3"],
                    help="fp32: CUDA-core parity mode (1e-4); x3: split-bf16 products on the tensor cores, fp32-grade (1e-4); "
                         "bf16: tcgen05 path (1e-2) with the device-side tie policy")
parser.add_argument("--checkpoint", default="./saved_checkpoints/mnist/checkpoint.pth.tar")
parser.add_argument("--no-write", action="store_true")


def conv(inp_chl, out_chl, ker_size=3, stride=1, padding=1):
    return nn.Sequential(nn.Conv2d(inp_chl, out_chl, ker_size, stride=stride, padding=padding),
                         nn.BatchNorm2d(out_chl), nn.ReLU(True))


class Classification_Net(EngineModule):
    input_hw = (28, 28)
    precision = "fp32"

    def __init__(self):
        super().__init__()
        self.conv1 = conv(1, 32)
        self.conv2 = conv(32, 32)
        self.conv3 = conv(32, 64, stride=2)
        self.conv4 = conv(64, 64)
        self.conv5 = conv(64, 128, stride=2)
        self.conv6 = nn.Conv2d(128, 128, 3, padding=1)
        self.fc1 = nn.Linear(128, 10)
        self._register_engine_hooks()

    def forward(self, x):
        pred0 = super().forward(x)
        net = self.engine()
        n = min(int(x.shape[0]), net.max_batch)     # taps hold the last micro-batch
        return net.read_tap("x0", n), net.read_tap("x1", n), net.read_tap("x2", n), pred0


def eval_superpixel(args):
    model = Classification_Net()
    if os.path.isfile(args.checkpoint):
        model.load_state_dict(torch.load(args.checkpoint, map_location="cpu", weights_only=False)["model"])   # :157-158
    model.eval()
    model.configure_engine(precision=args.precision, max_batch=256)
    image = np.load(args.image).astype(np.float32) if args.image else synthetic.synthetic_image("mnist")
    x0, x1, x2, pred0 = model(torch.from_numpy(image)[None].cuda())
    pred = int(pred0.argmax(1)[0])
    target = pred if args.target is None else args.target
    res = run_generator("mnist", model.engine(), image, target, args.num_mask_samples, args.mask_seed,
                        precision=args.precision, max_batch=256, mask_dir=None if args.no_write else "./masks")
    return res["correct_pred_count"], res["wrong_pred_count"]


if __name__ == "__main__":
    args = parser.parse_args()
    torch.manual_seed(args.seed)
    eval_superpixel(args)
