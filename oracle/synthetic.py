"""Deterministic synthetic inputs.  Oracle side: test infrastructure only."""
from __future__ import annotations

import numpy as np
import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # generate_gp_training_data_imagenet.py:590-591
IMAGENET_STD = (0.229, 0.224, 0.225)


def synthetic_image(kind: str, seed: int = 1234) -> np.ndarray:
    """C x H x W fp32 image in the space the reference's loader would deliver.

    imagenet: Normalize(mean,std) of u~U[0,1)  (imagenet :590-600)
    cifar   : (u-0.5)/0.5                      (generate_gp_training_data_cifar.py:52-54)
    mnist   : u (ToTensor only)                (generate_gp_training_data_mnist.py:59-62)
    """
    g = torch.Generator().manual_seed(seed)
    if kind == "imagenet":
        u = torch.rand(3, 224, 224, generator=g)
        mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
        std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
        return ((u - mean) / std).numpy().astype(np.float32)
    if kind == "cifar":
        u = torch.rand(3, 32, 32, generator=g)
        return ((u - 0.5) / 0.5).numpy().astype(np.float32)
    if kind == "mnist":
        return torch.rand(1, 28, 28, generator=g).numpy().astype(np.float32)
    raise ValueError(kind)


def voronoi_labels(H: int, W: int, S: int, seed: int = 7) -> np.ndarray:
    """Seeded Voronoi partition with contiguous labels 0..S-1 (int64 like felzenszwalb's output,
    imagenet :183).  Every label is guaranteed to own at least its site pixel."""
    rng = np.random.RandomState(seed)
    flat = rng.choice(H * W, size=S, replace=False)
    sy, sx = np.divmod(flat, W)
    yy, xx = np.mgrid[0:H, 0:W]
    d = (yy[None] - sy[:, None, None]) ** 2 + (xx[None] - sx[:, None, None]) ** 2
    lab = np.argmin(d, axis=0).astype(np.int64)
    lab[sy, sx] = np.arange(S)
    return lab
