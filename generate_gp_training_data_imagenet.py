"""Drop-in for the reference's generate_gp_training_data_imagenet.py (hot loop `validate()` :152-273).

    python generate_gp_training_data_imagenet.py -a resnet101 DIR [--num_mask_samples N] [--eval_img_index I]

The reference segments one validation image into superpixels, and for each of N random masks multiplies the image
by the mask, runs a batch-1 forward, checks top-1 and writes ./masks/mask_{i}_{0|1}.png (:221-266).  Here all N
masks are synthesised and scored on the B200 in micro-batches (masks sharded over ranks under torchrun) and the
PNGs are written afterwards.  Without a dataset / pretrained weights (no network in this image) the image is
synthetic and the torchvision model is seeded random-init: pass --image PATH / --weights PATH to use real ones.
Extra flags: --synthetic, --mask-seed, --subset (the commented variant :231), --precision {bf16,fp32}, --no-write."""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch
import torchvision.models as models

from network_interpretation_imagenet_b200 import synthetic
from network_interpretation_imagenet_b200.pipeline import run_generator

model_names = sorted(name for name in models.__dict__
                     if name.islower() and not name.startswith("__") and callable(models.__dict__[name]))

parser = argparse.ArgumentParser(description="PyTorch ImageNet perturbation-data generator (B200 engine)")
parser.add_argument("data", metavar="DIR", nargs="?", default=None, help="path to dataset (ImageFolder root with val/)")
parser.add_argument("--arch", "-a", metavar="ARCH", default="resnet18", choices=model_names)
parser.add_argument("-j", "--workers", default=4, type=int)
parser.add_argument("-b", "--batch-size", default=256, type=int, help="micro-batch of masks per forward")
parser.add_argument("--pretrained", dest="pretrained", action="store_true")
parser.add_argument("--world-size", default=1, type=int)
parser.add_argument("--dist-url", default="tcp://224.66.41.62:23456", type=str)
parser.add_argument("--dist-backend", default="gloo", type=str)
parser.add_argument("--eval_img_index", default=1600, type=int)
parser.add_argument("--num_mask_samples", default=100, type=int)
parser.add_argument("--synthetic", action="store_true", help="seeded synthetic image + label map + weights")
parser.add_argument("--image", default=None, help="path of one RGB image to explain instead of DIR/val[index]")
parser.add_argument("--weights", default=None, help="state_dict file for the torchvision model")
parser.add_argument("--target", default=None, type=int, help="class index (default: the model's top-1 on the image)")
parser.add_argument("--mask-seed", default=0, type=int)
parser.add_argument("--subset", action="store_true", help="random k-subset masks (reference :231) instead of windows")
parser.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
parser.add_argument("--no-write", action="store_true", help="skip the ./masks PNG side-channel")


def load_image(args) -> np.ndarray:
    """Resize(256) / CenterCrop(224) / ToTensor / Normalize of :590-600 on one image."""
    if args.image is None and (args.synthetic or args.data is None):
        return synthetic.synthetic_image("imagenet")
    import torchvision.datasets as datasets
    import torchvision.transforms as transforms
    tf = transforms.Compose([transforms.Resize(256), transforms.CenterCrop(224), transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    if args.image is not None:
        from PIL import Image
        return tf(Image.open(args.image).convert("RGB")).numpy()
    ds = datasets.ImageFolder(os.path.join(args.data, "val"), tf)
    img, _ = ds[args.eval_img_index - 1]            # the reference counts images from 1 (:163-166)
    return img.numpy()


def validate(model, image, args):
    """The reference's validate(): returns correct_pred_count (:268) or 0 when the unmasked image is misclassified."""
    from network_interpretation_imagenet_b200.classifier import Classifier
    # the unmasked image decides `target` (:190-200): one image, scored by the fp32 lowering
    logits = Classifier.from_torch(model, (224, 224), precision="fp32", max_batch=1).forward(torch.from_numpy(image)[None].cuda())
    pred = int(logits.argmax(1)[0])
    target = pred if args.target is None else args.target
    if pred != target:
        print("wrong prediction")
        return 0
    # the torch module (not a lowered Classifier) goes in, so the engine can lower the fp32 copy its tie policy needs
    res = run_generator("imagenet_subset" if args.subset else "imagenet", model, image, target, args.num_mask_samples,
                        args.mask_seed, precision=args.precision, max_batch=args.batch_size,
                        mask_dir=None if args.no_write else "./masks")
    return res["correct_pred_count"]


def main():
    args = parser.parse_args()
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    if args.weights:
        model = models.__dict__[args.arch](weights=None)
        model.load_state_dict(torch.load(args.weights, map_location="cpu"))
        model.eval()
    elif args.pretrained and not args.synthetic:
        model = models.__dict__[args.arch](pretrained=True).eval()     # needs the network, like the reference :579
    else:
        print("=> creating model '{}' (seeded random init: no pretrained weights available offline)".format(args.arch))
        model = synthetic.build_imagenet_model(args.arch)
    image = load_image(args)
    correct = validate(model, image, args)
    print("correct_pred_count", correct)


if __name__ == "__main__":
    main()
