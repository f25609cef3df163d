"""Host mirror of the pixel-coordinate GP regression of gp_regression.py over libnib.so.

`GridGPRegression` stands where the reference's `GPRegressionModel` (ExactGP + GridInterpolationKernel(RBF, grid_size=30,
grid_bounds=[(0, n), (0, n)]) + outputscale, gp_regression.py:160-176) and its prediction loop (:244-261) stand:
`fit(train_x, train_y)` then `predict(test_x)` -> (mean, variance) for every query pixel.  Arithmetic in csrc/ski.cu + gp.cu:

    K = L L^T (grid Gram, G = grid_size^2)        A = W^T W, b = W^T (y - c)   (sparse accumulation over the n pixels)
    B = s2 I + L^T A L = C C^T                    mean_U = L B^-1 L^T b         var(x*) = s2 |C^-1 L^T w*|^2 (+ s2)

i.e. the exact posterior of the SKI model in its inducing-weight form: two G x G Cholesky factorizations instead of an
n x n solve (n up to 224^2 = 50 176), and well conditioned whatever the length scale (cond(B) <= 1 + |A| |K| / s2).
`heatmap_from_masks` replaces the O(N * 224^2) Python dictionary loop of prepare_training_data (:63-104).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def heatmap_from_masks(masks_u8, labels, on_value: int = 255, device="cuda") -> torch.Tensor:
    """H[p] = sum_i labels[i] * [masks[i][p] == on_value]  (gp_regression.py:82-94).  masks [N,H,W] uint8 (numpy or
    CUDA tensor), labels [N].  Returns fp32 [H,W] on the device."""
    lib = _lib.load()
    m = torch.as_tensor(masks_u8).to(device=device, dtype=torch.uint8).contiguous()
    N, H, W = (int(v) for v in m.shape)
    y = torch.as_tensor(np.asarray(labels, dtype=np.float32)).to(device)
    heat = torch.zeros(H, W, dtype=torch.float32, device=device)
    _lib.check(lib.nib_heatmap_pixels(m.data_ptr(), y.data_ptr(), N, H * W, int(on_value), heat.data_ptr(),
                                      _lib.stream_handle()), "nib_heatmap_pixels")
    return heat


class GridGPRegression:
    def __init__(self, grid_size: int = 30, grid_bounds=((0.0, 224.0), (0.0, 224.0)), length_scale: float = 1.0,
                 outputscale: float = 1.0, noise: float = 1.0, const_mean: float = 0.0, jitter: float = 1e-8, device="cuda"):
        (lo0, hi0), (lo1, hi1) = grid_bounds
        if (lo0, hi0) != (lo1, hi1):
            raise ValueError("both dimensions must share their bounds (gp_regression.py:168 uses [(0, n), (0, n)])")
        self.gs = int(grid_size)
        d = (hi0 - lo0) / (self.gs - 2)                 # one spacing of margin on either side (cubic stencil)
        self.g0 = float(lo0 - d)
        self.h = float((hi0 - lo0 + 2 * d) / (self.gs - 1))
        self.ell, self.os, self.noise, self.c, self.jitter = float(length_scale), float(outputscale), float(noise), float(const_mean), float(jitter)
        self.device = torch.device(device)
        self.lib = _lib.load()

    def _dev64(self, a):
        if torch.is_tensor(a):
            return a.detach().to(device=self.device, dtype=torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)

    def _chol(self, M, what):
        G = M.shape[0]
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.nib_gp_cholesky(M.data_ptr(), G, G, info.data_ptr(), _lib.stream_handle()), "nib_gp_cholesky")
        if int(info.item()) != 0:
            raise np.linalg.LinAlgError(f"{what} is not positive definite (pivot {int(info.item())})")

    def fit(self, X, y):
        lib, st, dev = self.lib, _lib.stream_handle(), self.device
        X, y = self._dev64(X), self._dev64(y).reshape(-1)
        n, gs, G = int(X.shape[0]), self.gs, self.gs * self.gs
        if X.shape[1] != 2 or y.shape[0] != n:
            raise ValueError("train_x must be [n, 2] pixel coordinates and train_y [n]")
        c = self.g0 + self.h * torch.arange(gs, dtype=torch.float64, device=dev)
        U = torch.stack(torch.meshgrid(c, c, indexing="ij"), -1).reshape(G, 2).contiguous()
        L = torch.empty(G, G, dtype=torch.float64, device=dev)
        _lib.check(lib.nib_gp_gram_rbf(U.data_ptr(), G, U.data_ptr(), G, 2, self.ell, 0.0, 1, L.data_ptr(), G, st), "nib_gp_gram_rbf")
        L.mul_(self.os)
        L.diagonal().add_(self.jitter * self.os)
        self._chol(L, "the grid kernel K_UU")
        L.tril_()
        A = torch.zeros(G, G, dtype=torch.float64, device=dev)
        b = torch.zeros(G, dtype=torch.float64, device=dev)
        _lib.check(lib.nib_ski_accumulate(X.data_ptr(), y.data_ptr(), n, self.g0, self.h, gs, self.c, A.data_ptr(),
                                          b.data_ptr(), st), "nib_ski_accumulate")
        T = torch.zeros(G, G, dtype=torch.float64, device=dev)
        _lib.check(lib.nib_gp_dgemm_sub(A.data_ptr(), L.data_ptr(), T.data_ptr(), G, G, G, st), "nib_gp_dgemm_sub")   # T = -A L
        LT = L.t().contiguous()
        B = torch.zeros(G, G, dtype=torch.float64, device=dev)
        B.diagonal().fill_(self.noise)
        _lib.check(lib.nib_gp_dgemm_sub(LT.data_ptr(), T.data_ptr(), B.data_ptr(), G, G, G, st), "nib_gp_dgemm_sub")  # B = s2 I + L^T A L
        self._chol(B, "s2 I + L^T W^T W L")
        w = torch.empty(G, dtype=torch.float64, device=dev)
        _lib.check(lib.nib_gp_gemv_t(L.data_ptr(), G, G, G, b.data_ptr(), w.data_ptr(), st), "nib_gp_gemv_t")         # L^T b
        for trans in (0, 1):
            _lib.check(lib.nib_gp_trsm(B.data_ptr(), G, G, w.data_ptr(), 1, 1, trans, st), "nib_gp_trsm")
        self.mean_u = torch.empty(G, dtype=torch.float64, device=dev)
        _lib.check(lib.nib_gp_gemv_t(LT.data_ptr(), G, G, G, w.data_ptr(), self.mean_u.data_ptr(), st), "nib_gp_gemv_t")  # L w
        self.Gm = LT                                                                                                       # C^-1 L^T
        _lib.check(lib.nib_gp_trsm(B.data_ptr(), G, G, self.Gm.data_ptr(), G, G, 0, st), "nib_gp_trsm")
        self.n = n
        return self

    def predict(self, Xq, return_var: bool = True, likelihood: bool = True):
        """Posterior mean (and variance; with the Gaussian likelihood's noise when `likelihood`, as `likelihood(model(x))`
        at gp_regression.py:254) at pixel coordinates Xq [m, 2].  fp64 CUDA tensors."""
        Xq = self._dev64(Xq)
        m = int(Xq.shape[0])
        mean = torch.empty(m, dtype=torch.float64, device=self.device)
        var = torch.empty(m, dtype=torch.float64, device=self.device) if return_var else None
        _lib.check(self.lib.nib_ski_predict(Xq.data_ptr(), m, self.g0, self.h, self.gs, self.c, self.mean_u.data_ptr(),
                                            self.Gm.data_ptr(), self.noise, int(likelihood), mean.data_ptr(),
                                            var.data_ptr() if var is not None else None, _lib.stream_handle()), "nib_ski_predict")
        return (mean, var) if return_var else mean
