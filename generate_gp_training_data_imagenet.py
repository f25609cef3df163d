"""Drop-in for the reference's generate_gp_training_data_imagenet.py (hot loop `validate()` :152-273).

    python generate_gp_training_data_imagenet.py -a resnet101 DIR [--num_mask_samples N] [--eval_img_index I]

The reference segments one validation image into superpixels, and for each of N random masks multiplies the image
by the mask, runs a batch-1 forward, checks top-1 and writes ./masks/mask_{i}_{0|1}.png (:221-266).  Here all N
masks are synthesised and scored on the B200 in micro-batches (masks sharded over ranks under torchrun) and the
PNGs are written afterwards.  Without a dataset / pretrained weights (no network in this image) the image is
synthetic and the torchvision model is seeded random-init: pass --image PATH / --weights PATH to use real ones.
Extra flags: --synthetic, --mask-seed, --subset (the commented variant :231), --precision {bf16,fp32}, --no-write,
--mask-on-img (also write ./mask_on_img/), --validate-mask (run the threshold search of validate_mask, :334-488, on the
masks just scored: get_pixel_sorted_mask_label / plot_summed_heatmap / generate_new_mask keep their names)."""
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
parser.add_argument("--mask-on-img", action="store_true", help="also write ./mask_on_img/masked_imgs_{i}_{label}.png")
parser.add_argument("--validate-mask", action="store_true", help="threshold binary search on the summed-label heat map")


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
                        mask_dir=None if args.no_write else "./masks",
                        mask_on_img_dir="./mask_on_img" if (args.mask_on_img and not args.no_write) else None)
    validate.last = res
    return res["correct_pred_count"]


def get_pixel_sorted_mask_label(res=None):
    """Reference :490-515: dict_pixel[(j, k)] = sum of the labels of the masks whose pixel (j, k) is 255.  Returned as the
    dense device heat map plus the per-superpixel values / coverage it is constant on (from the masks just scored when
    `res` is given, else from the ./masks PNGs like the reference)."""
    from network_interpretation_imagenet_b200 import localize as loc
    if res is None:
        res = getattr(validate, "last", None)
    if res is not None:
        eng = res["engine"]
        labels = res["labels"].astype(np.float32)
        print("%d samples, the corrrect prediction number: %d " % (len(labels), int(labels.sum())))
        heat = eng.synth.heatmap(res["bits"], labels)
        wseg, cover = loc.segment_weights(eng.synth, res["bits"], labels)
        return {"heat": heat, "wseg": wseg, "cover": cover}
    import gp_regression
    gp_regression.n = 224
    heat, covered = gp_regression.summed_label_heatmap('./masks')
    return {"heat": heat, "covered": covered}


def generate_new_mask(dict_pixel, mask_threshold):
    """Reference :549-565 on the device heat map (uncovered pixels hold 0 and stay 0)."""
    import utils
    return utils.generate_new_mask(dict_pixel["heat"], mask_threshold)


def plot_summed_heatmap(val_img_index, org_img, label, dict_pixel):
    """Reference :517-546: the summed-label heat map as an 8-bit JET image (written to ./result_imgs instead of pyplot)."""
    import cv2
    from network_interpretation_imagenet_b200 import localize as loc
    gray = loc.heat_to_u8(dict_pixel["heat"]).cpu().numpy()
    os.makedirs("result_imgs", exist_ok=True)
    cv2.imwrite("result_imgs/index_{}_label_{}.png".format(val_img_index, label), cv2.applyColorMap(gray, cv2.COLORMAP_JET))
    return gray


def validate_mask(model, image, args, res=None):
    """Reference :334-488: binary search over the sorted distinct heat values for the largest threshold whose mask still
    gets the right top-1 while the next one does not.  All candidate masks are scored in one batch; the reference's probe
    sequence is replayed on the results (network_interpretation_imagenet_b200/localize.py)."""
    from network_interpretation_imagenet_b200 import localize as loc
    res = res if res is not None else validate.last
    eng = res["engine"]
    dict_pixel = get_pixel_sorted_mask_label(res)
    plot_summed_heatmap(args.eval_img_index, None, eng.target, dict_pixel)
    out = loc.threshold_search(eng, res["bits"], res["labels"].astype(np.float32), verbose=True)
    print("sorted_dict_values_set")
    print(out["values"])
    print("masked label threshold")
    print(out["threshold"])
    return out["threshold"]


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
    if args.validate_mask and getattr(validate, "last", None) is not None:
        validate_mask(model, image, args)


if __name__ == "__main__":
    main()
