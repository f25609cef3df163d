"""Mask semantics of the reference, restated in numpy.  Test infrastructure only.

Every function mirrors the cited reference lines operation by operation (same dtype promotions,
same in-place order), because the GPU path must match these bit for bit.
"""
from __future__ import annotations

import random as _random

import numpy as np


# ---- a1: per-image prep ------------------------------------------------------------------------
def prep_minmax_u8(org_img: np.ndarray):
    """generate_gp_training_data_mnist.py:169-177 / generate_gp_training_data_cifar.py:275-283.

    `img = org_img.transpose(1,2,0)` is a view, so the three in-place ops rescale `org_img` itself to
    [0,255]; `astype(np.uint8)` truncates.  Returns (org_img_rescaled (C,H,W) fp32, img_u8 (H,W,C))."""
    org = np.array(org_img, dtype=np.float32, copy=True)
    img = org.transpose(1, 2, 0)
    img -= img.min()
    img /= img.max()
    img *= 255
    return org, img.astype(np.uint8)


def normalize_image(x: np.ndarray) -> np.ndarray:
    """utils.py:92-94:  x.astype(np.float32); np.multiply(x, 1.0/255.0)."""
    x = x.astype(np.float32)
    return np.multiply(x, 1.0 / 255.0)


# ---- a3: selection draws (python `random`, module-global in the reference) -----------------------
def make_rng(seed: int) -> _random.Random:
    """random.Random(seed) yields exactly the sequence of `random.seed(seed)` + module functions."""
    return _random.Random(seed)


def draw_window(rng: _random.Random, u: np.ndarray):
    """Mode A, generate_gp_training_data_imagenet.py:223-230: k = int(0.4*S); f = randint(1, S-k);
    sel = np.unique(segments)[f:f+k]."""
    S = len(u)
    k = int(0.4 * S)
    f = rng.randint(1, S - k)
    return [int(v) for v in u[f:f + k]]


def window_selection(u: np.ndarray, first_index: int):
    """Mode A', bayesian_active_learning_imagenet.py:173-185: window at a BO-chosen firstIndex."""
    S = len(u)
    k = int(0.4 * S)
    return [int(v) for v in u[first_index:first_index + k]]


def draw_subset_mnist(rng: _random.Random, u: np.ndarray, k: int = 1):
    """Mode B (MNIST), generate_gp_training_data_mnist.py:208-215: a randint is drawn and discarded
    (:211), then sample(range(u[0], u[-1]), k) — `range` excludes the last label."""
    S = len(u)
    rng.randint(1, S - k)
    return rng.sample(range(int(u[0]), int(u[-1])), k)


def draw_subset_cifar(rng: _random.Random, u: np.ndarray, k: int = 5):
    """Mode B (CIFAR), generate_gp_training_data_cifar.py:308."""
    return rng.sample(range(int(u[0]), int(u[-1])), k)


def draw_subset_keep(rng: _random.Random, u: np.ndarray, k: int | None = None):
    """Mode B', the commented variant generate_gp_training_data_imagenet.py:231:
    sample(range(u[0], u[-1]), num_conse_superpixels) with num = int(0.4*S)."""
    S = len(u)
    if k is None:
        k = int(0.4 * S)
    return rng.sample(range(int(u[0]), int(u[-1])), k)


def selection_bits(selections, S: int) -> np.ndarray:
    """Selection lists -> [N, ceil(S/64)] uint64 bit-vectors z (bit s of word s//64)."""
    words = (S + 63) // 64
    out = np.zeros((len(selections), words), dtype=np.uint64)
    for n, sel in enumerate(selections):
        for s in sel:
            out[n, s // 64] |= np.uint64(1) << np.uint64(s % 64)
    return out


# ---- a4: pixel masks -------------------------------------------------------------------------------
def pixel_mask_keep(segments: np.ndarray, sel) -> np.ndarray:
    """imagenet :233-237: zeros(uint8); mask[segments == segVal] = 1."""
    mask = np.zeros(segments.shape[:2], dtype="uint8")
    for segVal in sel:
        mask[segments == segVal] = 1
    return mask


def pixel_mask_remove(segments: np.ndarray, sel) -> np.ndarray:
    """mnist :216-222 / cifar :310-313: fill(255); mask[segments == segVal] = 0."""
    mask = np.zeros(segments.shape[:2], dtype="uint8")
    mask.fill(255)
    for segVal in sel:
        mask[segments == segVal] = 0
    return mask


# ---- a5: apply + normalise -------------------------------------------------------------------------
def apply_keep(x_norm: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """imagenet :240, BO :187: `input[0].numpy().copy() * mask` (fp32 * uint8 -> fp32; x<0 gives -0.0)."""
    return x_norm.copy() * mask


def apply_remove_minmax(org_img: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """cifar :316-321 (MNIST :225-242 does the same through the alias `pic`):
    masked = org*mask; masked -= min; masked /= max; masked *= 255; normalize_image(masked)."""
    masked_img = org_img * mask
    masked_img -= masked_img.min()
    masked_img /= masked_img.max()
    masked_img *= 255
    return normalize_image(masked_img)


def masked_batch(x: np.ndarray, segments: np.ndarray, selections, mode: str) -> np.ndarray:
    """N masked classifier inputs, one reference loop iteration each. mode in {'keep','remove'}."""
    out = np.empty((len(selections),) + x.shape, dtype=np.float32)
    with np.errstate(invalid="ignore", divide="ignore"):
        for i, sel in enumerate(selections):
            if mode == "keep":
                out[i] = apply_keep(x, pixel_mask_keep(segments, sel))
            else:
                out[i] = apply_remove_minmax(x, pixel_mask_remove(segments, sel))
    return out


def heatmap(segments: np.ndarray, selections, labels) -> np.ndarray:
    """gp_regression.py:82-94 / gp_superpixel_data_imagenet.py:322-323: H = sum_i label_i * mask_i
    over keep-mode masks (pixels inside the kept superpixels accumulate the mask's label)."""
    H = np.zeros(segments.shape[:2], dtype=np.float64)
    for sel, y in zip(selections, labels):
        H += float(y) * pixel_mask_keep(segments, sel)
    return H
