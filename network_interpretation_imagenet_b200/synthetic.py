"""Deterministic synthetic inputs for `--synthetic` runs and bench.py: there is no network for
ImageNet images or pretrained weights, so the image, the superpixel label map and the classifier weights are
seeded.  Host-side setup only — nothing here is on the timed path."""
from __future__ import annotations

import numpy as np
import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # generate_gp_training_data_imagenet.py:590-591
IMAGENET_STD = (0.229, 0.224, 0.225)


def synthetic_image(kind: str, seed: int = 1234) -> np.ndarray:
    """C x H x W fp32 image in the space the reference's loader delivers (imagenet :590-600, cifar :52-54, mnist :59-62)."""
    g = torch.Generator().manual_seed(seed)
    if kind == "imagenet":
        u = torch.rand(3, 224, 224, generator=g)
        mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
        std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
        return ((u - mean) / std).numpy().astype(np.float32)
    if kind == "cifar":
        return ((torch.rand(3, 32, 32, generator=g) - 0.5) / 0.5).numpy().astype(np.float32)
    if kind == "mnist":
        return torch.rand(1, 28, 28, generator=g).numpy().astype(np.float32)
    raise ValueError(kind)


def voronoi_labels(H: int, W: int, S: int, seed: int = 7) -> np.ndarray:
    """Seeded Voronoi partition, contiguous labels 0..S-1 (stand-in for felzenszwalb, imagenet :183)."""
    rng = np.random.RandomState(seed)
    flat = rng.choice(H * W, size=S, replace=False)
    sy, sx = np.divmod(flat, W)
    yy, xx = np.mgrid[0:H, 0:W]
    d = (yy[None] - sy[:, None, None]) ** 2 + (xx[None] - sx[:, None, None]) ** 2
    lab = np.argmin(d, axis=0).astype(np.int64)
    lab[sy, sx] = np.arange(S)
    return lab


def randomize_bn(model: torch.nn.Module, seed: int) -> None:
    """Seeded, non-trivial BatchNorm statistics (random-init nets have BN = identity, which would make the fold a
    no-op and the logits nearly input-independent).  The last BN of every residual block is damped so the
    residual stream of a 101-layer random net stays O(1)."""
    g = torch.Generator().manual_seed(seed)
    for name, m in model.named_modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = 0.8 + 0.4 * torch.rand(m.num_features, generator=g)
            m.bias.data = 0.1 * torch.randn(m.num_features, generator=g)
            m.running_mean.data = 0.1 * torch.randn(m.num_features, generator=g)
            m.running_var.data = 0.8 + 0.4 * torch.rand(m.num_features, generator=g)
            if name.endswith("bn3"):   # torchvision Bottleneck's last BN
                m.weight.data *= 0.25


def build_imagenet_model(arch: str = "resnet101", seed: int = 0, randomize: bool = True) -> torch.nn.Module:
    """`models.__dict__[arch](pretrained=True)` of imagenet :579 with seeded random weights (no network here)."""
    import torchvision.models as tvm

    torch.manual_seed(seed)
    model = tvm.__dict__[arch](weights=None)
    if randomize:
        randomize_bn(model, seed + 1)
    return model.eval()
