"""Pixel-coordinate GP regression of gp_regression.py, restated in numpy.  Test infrastructure only.

The reference builds `GPRegressionModel` = gpytorch ExactGP + GridInterpolationKernel(RBFKernel, grid_size=30,
grid_bounds=[(0, n), (0, n)]) scaled by exp(log_outputscale), ConstantMean, GaussianLikelihood (gp_regression.py:160-176,
:405-408) and evaluates `likelihood(model(test_x))` on all n x n pixel coordinates (:244-261).  gpytorch is absent from
this image and unpinned by the reference (pre-0.1 API: `gpytorch.random_variables`, `log_lengthscale_bounds=`):
**parity unpinned**.  This file states the model by its DEFINITION — the dense n x n SKI covariance W K_UU W^T + s2 I solved
directly — so that the product's Woodbury / inducing-weight route (csrc/ski.cu, package ski.py) is checked against an
independent formulation, not against itself.

  grid      grid_size points per dimension from lo - d to hi + d, d = (hi - lo) / (grid_size - 2)  (one spacing of margin so
            the 4-point cubic stencil of any in-bounds point stays on the grid)
  W         cubic convolution interpolation (Keys 1981, a = -0.5), 4 x 4 = 16 non-zeros per row
  K_UU      outputscale * exp(-|u - u'|^2 / (2 l^2))
  defaults  l = outputscale = noise = 1, constant mean 0: the reference's training loop never steps its optimiser in the
            branch it runs (gp_regression.py:206-217 has no backward()/step()), so it evaluates the initial parameters.
"""
from __future__ import annotations

import numpy as np


def make_grid(lo: float, hi: float, grid_size: int):
    d = (hi - lo) / (grid_size - 2)
    g0 = lo - d
    h = (hi - lo + 2 * d) / (grid_size - 1)
    return g0, h


def keys(s: np.ndarray) -> np.ndarray:
    s = np.abs(s)
    return np.where(s <= 1.0, (1.5 * s - 2.5) * s * s + 1.0, np.where(s < 2.0, ((-0.5 * s + 2.5) * s - 4.0) * s + 2.0, 0.0))


def interp_matrix(X: np.ndarray, g0: float, h: float, gs: int) -> np.ndarray:
    """Dense W [n, gs*gs] (row-major grid: index = i0 * gs + i1)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    W = np.zeros((n, gs * gs))
    u = (X - g0) / h
    base = np.floor(u).astype(int) - 1
    for a in range(4):
        for b in range(4):
            i0 = np.clip(base[:, 0] + a, 0, gs - 1)
            i1 = np.clip(base[:, 1] + b, 0, gs - 1)
            w = keys(u[:, 0] - (base[:, 0] + a)) * keys(u[:, 1] - (base[:, 1] + b))
            np.add.at(W, (np.arange(n), i0 * gs + i1), w)
    return W


def grid_kernel(g0: float, h: float, gs: int, length_scale: float, outputscale: float) -> np.ndarray:
    c = g0 + h * np.arange(gs)
    U = np.stack(np.meshgrid(c, c, indexing="ij"), -1).reshape(-1, 2)
    d2 = ((U[:, None, :] - U[None, :, :]) ** 2).sum(-1)
    return outputscale * np.exp(-0.5 * d2 / (length_scale * length_scale))


def ski_posterior(X, y, Xq, bounds=(0.0, 224.0), grid_size=30, length_scale=1.0, outputscale=1.0, noise=1.0,
                  const_mean=0.0, likelihood=True):
    """Posterior mean / variance at Xq by the dense definition (n must be small enough for an n x n solve)."""
    g0, h = make_grid(bounds[0], bounds[1], grid_size)
    K = grid_kernel(g0, h, grid_size, length_scale, outputscale)
    W = interp_matrix(X, g0, h, grid_size)
    Wq = interp_matrix(Xq, g0, h, grid_size)
    Kxx = W @ K @ W.T + noise * np.eye(W.shape[0])
    Kqx = Wq @ K @ W.T
    sol = np.linalg.solve(Kxx, np.asarray(y, dtype=np.float64) - const_mean)
    mean = const_mean + Kqx @ sol
    var = np.einsum("ij,ij->i", Wq @ K, Wq) - np.einsum("ij,ji->i", Kqx, np.linalg.solve(Kxx, Kqx.T))
    return mean, var + (noise if likelihood else 0.0)


def heatmap_from_masks(masks_u8: np.ndarray, labels: np.ndarray) -> np.ndarray:
    """gp_regression.py:82-104: result_gray_img[p] = sum of mask labels over the masks whose pixel p is 255."""
    N = masks_u8.shape[0]
    return ((masks_u8.reshape(N, -1) == 255) * np.asarray(labels, dtype=np.float64)[:, None]).sum(0).reshape(masks_u8.shape[1:])
