"""Host mirror of the variational GP classifier of gp_classification.py over libnib.so.

`GridVariationalGPClassifier` stands where the reference's `GPClassificationModel` (gpytorch GridInducingVariationalGP,
grid_size = 10 over [0, n]^2, near-zero ConstantMean, RBF x exp(log_outputscale); gp_classification.py:139-156) with its
BernoulliLikelihood and the 30-step Adam(lr = 0.1) loop on the negative variational ELBO (:160-217) stand; `predict_proba`
is `likelihood(model(x)).mean()` (:241-253).

Split of the work: everything that scales with the number of training pixels n (up to 224^2 = 50 176) - the expected
log-likelihood and its gradients w.r.t. the variational mean and covariance, and the predictive probabilities - runs in
csrc/ski.cu kernels (sparse 4 x 4 cubic stencils, quadrature, fp64 atomics into G / G x G arrays).  The KL term and the
Adam update touch only G x G matrices (G = grid_size^2 = 100): host numpy, O(G^3) = 1e6 flops per step.
gpytorch is absent and unpinned by the reference: parity unpinned; pinned instead to the dense definition in
oracle/gp_classification.py (tests/test_gpu_vgp.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class GridVariationalGPClassifier:
    def __init__(self, grid_size: int = 10, grid_bounds=((0.0, 224.0), (0.0, 224.0)), log_lengthscale: float = 0.0,
                 log_outputscale: float = 0.0, const_mean: float = 0.0, log_lengthscale_bounds=(-5.0, 6.0),
                 log_outputscale_bounds=(-5.0, 6.0), const_mean_bounds=(-1e-5, 1e-5), jitter: float = 1e-6, device="cuda"):
        (lo0, hi0), (lo1, hi1) = grid_bounds
        if (lo0, hi0) != (lo1, hi1):
            raise ValueError("both dimensions must share their bounds (gp_classification.py:141 uses [(0, n), (0, n)])")
        self.gs = int(grid_size)
        d = (hi0 - lo0) / (self.gs - 2)                 # one spacing of margin on either side (cubic stencil)
        self.g0 = float(lo0 - d)
        self.h = float((hi0 - lo0 + 2 * d) / (self.gs - 1))
        self.G = self.gs * self.gs
        self.log_lengthscale, self.log_outputscale, self.const_mean = float(log_lengthscale), float(log_outputscale), float(const_mean)
        self.bounds = {"log_lengthscale": log_lengthscale_bounds, "log_outputscale": log_outputscale_bounds,
                       "const_mean": const_mean_bounds}
        self.jitter = float(jitter)
        self.variational_mean = np.zeros(self.G)
        self.chol_variational_covar = np.eye(self.G)
        self.device = torch.device(device)
        self.lib = _lib.load()
        c = self.g0 + self.h * np.arange(self.gs)
        U = np.stack(np.meshgrid(c, c, indexing="ij"), -1).reshape(-1, 2)
        self._d2 = ((U[:, None, :] - U[None, :, :]) ** 2).sum(-1)
        self.history: list[dict] = []

    # gpytorch-module surface the reference's train() / eval_superpixels() call (gp_classification.py:163-164,:226-227)
    def cuda(self):
        return self

    def train(self):
        return self

    def eval(self):
        return self

    # -- state ------------------------------------------------------------------------------------------
    def state_dict(self):
        return {"log_lengthscale": self.log_lengthscale, "log_outputscale": self.log_outputscale, "const_mean": self.const_mean,
                "variational_mean": self.variational_mean.copy(), "chol_variational_covar": self.chol_variational_covar.copy(),
                "grid_size": self.gs}

    def load_state_dict(self, sd):
        if int(sd.get("grid_size", self.gs)) != self.gs:
            raise ValueError("checkpoint grid size does not match the model")
        self.log_lengthscale, self.log_outputscale = float(sd["log_lengthscale"]), float(sd["log_outputscale"])
        self.const_mean = float(sd["const_mean"])
        self.variational_mean = np.asarray(sd["variational_mean"], dtype=np.float64).copy()
        self.chol_variational_covar = np.asarray(sd["chol_variational_covar"], dtype=np.float64).copy()

    # -- pieces -----------------------------------------------------------------------------------------
    def _dev64(self, a):
        if torch.is_tensor(a):
            return a.detach().to(device=self.device, dtype=torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)

    def _prior(self):
        ell2, os_ = np.exp(2.0 * self.log_lengthscale), np.exp(self.log_outputscale)
        K0 = os_ * np.exp(-0.5 * self._d2 / ell2)
        K = K0.copy()
        K[np.diag_indices_from(K)] += self.jitter * os_
        return K0, K

    def data_term(self, X_d, y_d):
        """(sum_i E[log Phi(y_i f_i)], dE/dm [G], dE/dS [G,G], dE/dc) from the device kernel."""
        G, dev = self.G, self.device
        Ls = np.tril(self.chol_variational_covar)
        m_d = self._dev64(self.variational_mean)
        S_d = self._dev64(Ls @ Ls.T)
        ell = torch.empty(1, dtype=torch.float64, device=dev)
        gm = torch.empty(G, dtype=torch.float64, device=dev)
        gS = torch.empty(G, G, dtype=torch.float64, device=dev)
        gc = torch.empty(1, dtype=torch.float64, device=dev)
        _lib.check(self.lib.nib_vgp_loglik_grad(X_d.data_ptr(), y_d.data_ptr(), int(X_d.shape[0]), self.g0, self.h, self.gs,
                                                self.const_mean, m_d.data_ptr(), S_d.data_ptr(), ell.data_ptr(), gm.data_ptr(),
                                                gS.data_ptr(), gc.data_ptr(), _lib.stream_handle()), "nib_vgp_loglik_grad")
        return float(ell.item()), gm.cpu().numpy(), gS.cpu().numpy(), float(gc.item())

    def kl_term(self):
        """KL(q(u) || p(u)) and its gradients w.r.t. m, Ls, log_lengthscale, log_outputscale (G x G algebra on the host)."""
        m, Ls = self.variational_mean, np.tril(self.chol_variational_covar)
        K0, K = self._prior()
        S = Ls @ Ls.T
        Kinv = np.linalg.inv(K)
        G = self.G
        kl = 0.5 * (np.trace(Kinv @ S) + m @ Kinv @ m - G + np.linalg.slogdet(K)[1] - 2.0 * np.log(np.abs(np.diag(Ls))).sum())
        g_m = Kinv @ m
        g_Ls = np.tril(Kinv @ Ls - np.linalg.inv(Ls).T)
        Mk = Kinv - Kinv @ (S + np.outer(m, m)) @ Kinv          # dKL/dK = 1/2 Mk
        dK_dlogl = K0 * self._d2 / np.exp(2.0 * self.log_lengthscale)
        g_logl = 0.5 * np.sum(Mk * dK_dlogl)
        g_logos = 0.5 * np.sum(Mk * K)
        return kl, g_m, g_Ls, g_logl, g_logos

    def loss_and_grads(self, X_d, y_d):
        """loss = -(sum_i E_i - KL) / n  (VariationalMarginalLogLikelihood divides both terms by n_data) and its gradients."""
        n = int(X_d.shape[0])
        ell, gm_d, gS_d, gc_d = self.data_term(X_d, y_d)
        kl, gm_k, gLs_k, g_logl, g_logos = self.kl_term()
        Ls = np.tril(self.chol_variational_covar)
        g = {"variational_mean": -(gm_d - gm_k) / n,
             "chol_variational_covar": -(np.tril((gS_d + gS_d.T) @ Ls) - gLs_k) / n,
             "log_lengthscale": g_logl / n, "log_outputscale": g_logos / n, "const_mean": -gc_d / n}
        return -(ell - kl) / n, g, {"expected_loglik": ell, "kl": kl}

    # -- training (gp_classification.py:160-217) -----------------------------------------------------------
    def fit(self, train_x, train_y, num_training_iterations: int = 30, lr: float = 0.1, verbose: bool = True):
        X_d, y_d = self._dev64(train_x), self._dev64(train_y).reshape(-1)
        if X_d.dim() != 2 or X_d.shape[1] != 2 or y_d.shape[0] != X_d.shape[0]:
            raise ValueError("train_x must be [n, 2] pixel coordinates and train_y [n]")
        names = ["variational_mean", "chol_variational_covar", "log_lengthscale", "log_outputscale", "const_mean"]
        st = {k: (np.zeros_like(np.asarray(getattr(self, k), dtype=np.float64)), np.zeros_like(np.asarray(getattr(self, k), dtype=np.float64)))
              for k in names}
        b1, b2, eps = 0.9, 0.999, 1e-8                  # torch.optim.Adam defaults
        for it in range(1, num_training_iterations + 1):
            loss, g, parts = self.loss_and_grads(X_d, y_d)
            if verbose:
                print('Iter %d/%d - Loss: %.3f   log_lengthscale: %.3f' % (it, num_training_iterations, loss, self.log_lengthscale))
            self.history.append({"iter": it, "loss": loss, **parts})
            for k in names:
                m1, m2 = st[k]
                gk = np.asarray(g[k], dtype=np.float64)
                m1 = b1 * m1 + (1 - b1) * gk
                m2 = b2 * m2 + (1 - b2) * gk * gk
                st[k] = (m1, m2)
                step = lr * (m1 / (1 - b1 ** it)) / (np.sqrt(m2 / (1 - b2 ** it)) + eps)
                new = np.asarray(getattr(self, k), dtype=np.float64) - step
                if k in self.bounds:
                    new = np.clip(new, *self.bounds[k])
                if k == "chol_variational_covar":
                    new = np.tril(new)
                setattr(self, k, float(new) if new.ndim == 0 else new)
        return self

    # -- prediction (gp_classification.py:241-253) ----------------------------------------------------------
    def predict_proba(self, Xq, return_latent: bool = False):
        Xq = self._dev64(Xq)
        mq, dev = int(Xq.shape[0]), self.device
        Ls = np.tril(self.chol_variational_covar)
        m_d, S_d = self._dev64(self.variational_mean), self._dev64(Ls @ Ls.T)
        prob = torch.empty(mq, dtype=torch.float64, device=dev)
        mu = torch.empty(mq, dtype=torch.float64, device=dev) if return_latent else None
        var = torch.empty(mq, dtype=torch.float64, device=dev) if return_latent else None
        _lib.check(self.lib.nib_vgp_predict(Xq.data_ptr(), mq, self.g0, self.h, self.gs, self.const_mean, m_d.data_ptr(),
                                            S_d.data_ptr(), prob.data_ptr(), mu.data_ptr() if mu is not None else None,
                                            var.data_ptr() if var is not None else None, _lib.stream_handle()), "nib_vgp_predict")
        return (prob, mu, var) if return_latent else prob
