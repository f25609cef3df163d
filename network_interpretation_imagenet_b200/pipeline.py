"""Shared body of the drop-in entry scripts (generate_gp_training_data_{mnist,cifar,imagenet}.py,
bayesian_active_learning_imagenet.py): everything the reference inlines in its `validate()` / `eval_superpixel()`
hot loops, expressed once over the engine.  Also the on-disk side-channel the GP scripts read
(`./masks/mask_{i}_{label}.png`, gp_regression.py:63-75), written in batches off the device."""
from __future__ import annotations

import os
import shutil

import numpy as np
import torch

from .engine import PerturbationEngine
from .masks import KEEP_MUL, REMOVE_MINMAX, draw_selections, prep_minmax_u8, selection_bits


def felzenszwalb(image: np.ndarray, scale: float = 1.0, sigma: float = 0.8, min_size: int = 20) -> np.ndarray:
    """Drop-in for `skimage.segmentation.felzenszwalb(image, scale, sigma, min_size)` on a float image in [0,1]
    (H x W or H x W x C): libnib's native restatement of scikit-image's algorithm (csrc/segment.cu).  int64 labels H x W,
    contiguous 0..S-1 in raster order of first appearance."""
    import ctypes as C

    from . import _lib
    img = np.ascontiguousarray(np.atleast_3d(np.asarray(image, dtype=np.float64)))
    H, W, Cc = img.shape
    labels = np.empty((H, W), dtype=np.int32)
    S = C.c_int()
    _lib.check(_lib.load().nib_felzenszwalb(img.ctypes.data, H, W, Cc, float(scale), float(sigma), int(min_size),
                                            labels.ctypes.data, C.byref(S)), "nib_felzenszwalb")
    return labels.astype(np.int64)


def img_as_float(img_u8: np.ndarray) -> np.ndarray:
    """`skimage.util.img_as_float` of a uint8 image: float64, multiplied by 1/255 (scikit-image scales unsigned input
    with `np.multiply(image, 1. / imax_in)`, not a division)."""
    if img_u8.dtype != np.uint8:
        raise TypeError("img_as_float restated for uint8 input only (what the reference passes, imagenet :178,:183)")
    return np.multiply(img_u8, 1.0 / 255.0, dtype=np.float64)


def segment_image(img_u8_hwc: np.ndarray, min_size: int, S_fallback: int | None = None, seed: int = 7) -> np.ndarray:
    """`felzenszwalb(img_as_float(img), scale=100, sigma=0.5, min_size=...)` (imagenet :183, cifar :293, mnist :187),
    computed by the library's own implementation (scikit-image is not needed).  S_fallback/seed are accepted for
    backwards compatibility and ignored."""
    return felzenszwalb(img_as_float(np.asarray(img_u8_hwc)), scale=100, sigma=0.5, min_size=min_size)


def reset_dir(path: str) -> None:
    """mask_dir handling of imagenet :207-212: create, or wipe and re-create."""
    if os.path.exists(path):
        shutil.rmtree(path)
    os.makedirs(path)


def write_mask_pngs(mask_dir: str, pixel_masks_u8: np.ndarray, labels, scale: int, start: int = 0) -> None:
    """cv2.imwrite('./masks/mask_{i}_{label}.png', mask*scale)  (imagenet :260/:265 scale=255; mnist/cifar scale=1)."""
    import cv2
    for j, (m, lab) in enumerate(zip(pixel_masks_u8, labels)):
        cv2.imwrite(os.path.join(mask_dir, "mask_{}_{}.png".format(start + j, int(lab))), m * scale)


def run_generator(kind: str, model, image_chw: np.ndarray, target: int, n_masks: int, mask_seed: int,
                  precision: str = "bf16", max_batch: int = 256, segments: np.ndarray | None = None,
                  mask_dir: str | None = None, S_fallback: int | None = None, verbose: bool = True):
    """One image through the reference's generator loop.  kind: 'imagenet' (keep window, :221-266),
    'imagenet_subset' (the commented variant :231), 'mnist' (:203-269), 'cifar' (:307-342).
    Returns dict(correct_pred_count, wrong_pred_count, labels u8[N], target_prob f32[N], top1 i32[N], bits, segments)."""
    remove = kind in ("mnist", "cifar")
    if remove:
        d_img, u8 = prep_minmax_u8(image_chw)                     # a1, in place on the device
        u8 = u8.cpu().numpy()
    else:
        disp = image_chw.transpose(1, 2, 0).copy()                # imagenet :171-178 works on a copy
        disp -= disp.min(); disp /= disp.max(); disp *= 255
        u8 = disp.astype(np.uint8)
        d_img = image_chw
    if segments is None:
        defaults = {"imagenet": (50, 50), "imagenet_subset": (50, 50), "cifar": (10, 20), "mnist": (5, 16)}[kind]
        segments = segment_image(u8, defaults[0], S_fallback or defaults[1])
    u = np.unique(segments)
    if not np.array_equal(u, np.arange(len(u))):
        raise ValueError("segment labels must be contiguous 0..S-1")
    S = len(u)
    draw = {"imagenet": "window", "imagenet_subset": "subset_keep", "mnist": "mnist", "cifar": "cifar"}[kind]
    sels = draw_selections(draw, S, n_masks, seed=mask_seed)
    bits = selection_bits(sels, S)
    eng = PerturbationEngine(model, d_img, segments, target, mode=REMOVE_MINMAX if remove else KEEP_MUL,
                             precision=precision, max_batch=max_batch, S=S)
    out = eng.score_masks(bits)
    labels = out["correct"].to(torch.uint8).cpu().numpy()
    res = {"correct_pred_count": int(labels.sum()), "wrong_pred_count": int(n_masks - labels.sum()), "labels": labels,
           "target_prob": out["target_prob"].cpu().numpy(), "top1": out["top1"].cpu().numpy(), "bits": bits,
           "segments": segments, "selections": sels, "engine": eng}
    if mask_dir is not None and eng.rank == 0:
        reset_dir(mask_dir)
        for i in range(0, n_masks, 1024):
            _, pm = eng.synth.synth(bits[i:i + 1024], eng.mode, return_pixel_masks=True)
            write_mask_pngs(mask_dir, pm.cpu().numpy(), labels[i:i + 1024], 1 if remove else 255, start=i)
    if verbose and eng.rank == 0:
        print("correct_pred_count: ", res["correct_pred_count"])
        print("wrong_pred_count: ", res["wrong_pred_count"])
    return res
