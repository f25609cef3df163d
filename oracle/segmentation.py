"""Superpixel segmentation of the reference, restated in numpy.  Test infrastructure only.

The reference calls `felzenszwalb(img_as_float(img), scale=100, sigma=0.5, min_size=50)`
(generate_gp_training_data_imagenet.py:183; mnist :187 min_size=5; cifar :293 min_size=10;
bayesian_active_learning_imagenet.py:150,:263,:463).  That function lives in scikit-image
(`skimage/segmentation/_felzenszwalb.py` -> `_felzenszwalb_cy.pyx`), a dependency the reference does not pin
(requirements.txt:1-2 lists torch/torchvision only) and that is absent from this image: **parity unpinned** — this
file restates the published algorithm of scikit-image's implementation (Felzenszwalb & Huttenlocher 2004,
"Efficient graph-based image segmentation", as coded in scikit-image 0.14 ... 0.25, unchanged over that span):

  1. image as float64 in [0,1] (`img_as_float` of uint8 = x * (1/255), a multiplication), grey images get a trailing channel axis;
  2. `scale /= 255`; Gaussian blur `scipy.ndimage.gaussian_filter(image, sigma=[sigma, sigma, 0])`
     (mode 'reflect', truncate 4.0 -> radius int(4*sigma+0.5); axis 0 then axis 1);
  3. 8-connected grid graph, edge cost = Euclidean colour distance of the blurred pixels, edges listed as
     right, down, down-right, up-right blocks (row-major inside each block) and argsorted by cost;
  4. pass 1 over sorted edges: merge components a, b if cost < min(int[a] + scale/|a|, int[b] + scale/|b|);
     the merged component's internal cost becomes this edge's cost;
  5. pass 2 over the same edge order: merge if either component is smaller than min_size;
  6. labels = rank of each component's smallest pixel index (np.unique(..., return_inverse=True)): union keeps the
     smaller root, so labels are numbered in raster order of first appearance.

scikit-image argsorts with numpy's default (unstable) sort; ties between *positive* equal costs could therefore
order differently there.  Zero-cost ties always merge whatever the order.  Both this restatement and the product
use a stable sort (edge index breaks ties).
"""
from __future__ import annotations

import numpy as np


def img_as_float(img_u8: np.ndarray) -> np.ndarray:
    """skimage.util.img_as_float for uint8 input (`dtype._convert`: `np.multiply(image, 1. / imax_in, dtype=float64)`)."""
    assert img_u8.dtype == np.uint8
    return np.multiply(img_u8, 1.0 / 255.0, dtype=np.float64)


def gaussian_kernel1d(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius) with radius = int(truncate*sigma + 0.5)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def _correlate1d_reflect(a: np.ndarray, w: np.ndarray, axis: int) -> np.ndarray:
    """scipy.ndimage.correlate1d, mode='reflect' (d c b a | a b c d | d c b a), symmetric-kernel accumulation order
    of ni_filters.c: centre tap first, then pairs (x[l+j] + x[l-j]) * w from the outermost pair inwards."""
    r = len(w) // 2
    a = np.moveaxis(a, axis, 0)
    n = a.shape[0]
    idx = np.arange(-r, n + r)
    idx = np.where(idx < 0, -idx - 1, idx)
    idx = np.where(idx >= n, 2 * n - 1 - idx, idx)
    # reflect may need several folds for tiny axes; images here are always larger than the radius
    p = a[idx]
    out = p[r:r + n] * w[r]
    for j in range(-r, 0):
        out = out + (p[r + j:r + j + n] + p[r - j:r - j + n]) * w[j + r]
    return np.moveaxis(out, 0, axis)


def gaussian_blur(image_hwc: np.ndarray, sigma: float) -> np.ndarray:
    w = gaussian_kernel1d(sigma)
    return _correlate1d_reflect(_correlate1d_reflect(image_hwc, w, 0), w, 1)


def grid_edges(blurred: np.ndarray):
    """Edge list (a, b) and costs in scikit-image's block order: right, down, down-right, up-right."""
    H, W = blurred.shape[:2]
    im = blurred

    def cost(x, y):
        d = x - y
        return np.sqrt(np.sum(d * d, axis=-1))

    right_cost = cost(im[:, 1:], im[:, :W - 1])
    down_cost = cost(im[1:], im[:H - 1])
    dright_cost = cost(im[1:, 1:], im[:H - 1, :W - 1])
    uright_cost = cost(im[1:, :W - 1], im[:H - 1, 1:])
    costs = np.hstack([right_cost.ravel(), down_cost.ravel(), dright_cost.ravel(), uright_cost.ravel()]).astype(float)
    seg = np.arange(W * H, dtype=np.intp).reshape(H, W)
    right_edges = np.c_[seg[:, 1:].ravel(), seg[:, :W - 1].ravel()]
    down_edges = np.c_[seg[1:].ravel(), seg[:H - 1].ravel()]
    dright_edges = np.c_[seg[1:, 1:].ravel(), seg[:H - 1, :W - 1].ravel()]
    uright_edges = np.c_[seg[:H - 1, 1:].ravel(), seg[1:, :W - 1].ravel()]
    edges = np.vstack([right_edges, down_edges, dright_edges, uright_edges])
    return edges, costs


def felzenszwalb(image: np.ndarray, scale: float = 1.0, sigma: float = 0.8, min_size: int = 20) -> np.ndarray:
    """image: H x W or H x W x C, float in [0,1] (what `img_as_float(u8)` yields).  Returns int64 labels H x W."""
    image = np.atleast_3d(np.asarray(image, dtype=np.float64))
    H, W = image.shape[:2]
    scale = float(scale) / 255.0
    edges, costs = grid_edges(gaussian_blur(image, sigma))
    order = np.argsort(costs, kind="stable")
    edges, costs = edges[order].tolist(), costs[order].tolist()
    parent = list(range(H * W))
    size = [1] * (H * W)
    cint = [0.0] * (H * W)

    def find(n):
        root = n
        while parent[root] != root:
            root = parent[root]
        while parent[n] != root:            # path compression (does not change which pixels share a root)
            parent[n], n = root, parent[n]
        return root

    for (a, b), c in zip(edges, costs):
        s0, s1 = find(a), find(b)
        if s0 == s1:
            continue
        if c < min(cint[s0] + scale / size[s0], cint[s1] + scale / size[s1]):
            new, old = (s0, s1) if s0 < s1 else (s1, s0)      # join_trees keeps the smaller root
            parent[old] = new
            size[new] = size[s0] + size[s1]
            cint[new] = c
    for a, b in edges:
        s0, s1 = find(a), find(b)
        if s0 == s1:
            continue
        if size[s0] < min_size or size[s1] < min_size:
            new, old = (s0, s1) if s0 < s1 else (s1, s0)
            parent[old] = new
            size[new] = size[s0] + size[s1]
    roots = np.array([find(i) for i in range(H * W)], dtype=np.int64)
    return np.unique(roots, return_inverse=True)[1].reshape(H, W).astype(np.int64)
