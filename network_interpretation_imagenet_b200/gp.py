"""Stage 3 host mirror: Gaussian-process surrogate + Expected Improvement on the device.

Duck-types the estimator the reference builds at BayesianOptimization.py:154-159
(`GaussianProcessRegressor(kernel=RBF(), alpha=1e-5, n_restarts_optimizer=10, normalize_y=True)`):
`fit(X, y)` (:166) and `predict(X, return_std=True)` (:39) with numpy fp64 in/out, plus
`log_marginal_likelihood(theta, eval_gradient=True)` following sklearn/gaussian_process/_gpr.py:588-655.
The hyper-parameter search keeps scikit-learn's host control flow (scipy L-BFGS-B on theta = log l from
the initial l and `n_restarts_optimizer` log-uniform restarts, _gpr.py:298-339) while every O(n^2)/O(n^3)
step — Gram, Cholesky, solves, K^-1, posterior, EI — runs in libnib.so kernels.

X may be (a) selection bit-vectors [n, words] uint64 (the scaled-up spec: masks as inputs),
(b) a 0/1 matrix [n, S], converted to (a), or (c) any real matrix [n, d] (the reference's 1-D firstIndex GP).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from scipy.optimize import minimize

from . import _lib


def _is_bits(X) -> bool:
    return isinstance(X, np.ndarray) and X.dtype == np.uint64


def pack_bits(X01: np.ndarray) -> np.ndarray:
    X01 = np.asarray(X01)
    n, S = X01.shape
    words = (S + 63) // 64
    out = np.zeros((n, words), dtype=np.uint64)
    for s in range(S):
        out[:, s >> 6] |= (X01[:, s] != 0).astype(np.uint64) << np.uint64(s & 63)
    return out


class GaussianProcessRegressor:
    def __init__(self, alpha: float = 1e-5, normalize_y: bool = True, length_scale: float = 1.0,
                 length_scale_bounds=(1e-5, 1e5), n_restarts_optimizer: int = 0, optimizer: str | None = "fmin_l_bfgs_b",
                 random_state=None, kernel=None, device="cuda", query_chunk: int = 8192):
        # `kernel` is accepted for signature compatibility with gp_params passthrough (BayesianOptimization.py:150-151);
        # only RBF is implemented, its length_scale / bounds are read if present.
        if kernel is not None:
            length_scale = float(getattr(kernel, "length_scale", length_scale))
            length_scale_bounds = tuple(getattr(kernel, "length_scale_bounds", length_scale_bounds))
        self.alpha = float(alpha)
        self.normalize_y = normalize_y
        self.length_scale0 = float(length_scale)
        self.length_scale_bounds = length_scale_bounds
        self.n_restarts_optimizer = int(n_restarts_optimizer)
        self.optimizer = optimizer
        self.random_state = random_state
        self.device = torch.device(device)
        self.query_chunk = int(query_chunk)
        self.lib = _lib.load()
        self.length_scale_ = self.length_scale0
        self.n = 0
        self.stats = {"lml_evals": 0}

    # ---- helpers -------------------------------------------------------------------------------
    def _upload_X(self, X):
        X = np.asarray(X)
        if _is_bits(X):
            bits = np.ascontiguousarray(X)
            return "bits", torch.from_numpy(bits.view(np.int64)).to(self.device), bits.shape[1]
        if X.ndim == 2 and X.shape[1] > 1 and np.isin(X, (0, 1)).all():
            bits = pack_bits(X)
            return "bits", torch.from_numpy(bits.view(np.int64)).to(self.device), bits.shape[1]
        Xr = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
        return "real", torch.from_numpy(Xr).to(self.device), Xr.shape[1]

    def _gram(self, kind, A, na, B, nb, dim, ell, jitter, same, out):
        st = _lib.stream_handle()
        if kind == "bits":
            a_ptr, b_ptr = A.data_ptr(), (A.data_ptr() if same else B.data_ptr())
            _lib.check(self.lib.nib_gp_gram_binary(a_ptr, na, b_ptr, nb, dim, ell, jitter, out.data_ptr(),
                                                   out.stride(0), st), "nib_gp_gram_binary")
        else:
            _lib.check(self.lib.nib_gp_gram_rbf(A.data_ptr(), na, B.data_ptr(), nb, dim, ell, jitter, int(same),
                                                out.data_ptr(), out.stride(0), st), "nib_gp_gram_rbf")

    def _factor(self, ell: float):
        """K(+alpha I) -> L (in place), alpha_vec = K^-1 y.  Raises LinAlgError like _gpr.py:352-361."""
        n = self.n
        st = _lib.stream_handle()
        self._gram(self.kind, self.X_d, n, self.X_d, n, self.dim, ell, self.alpha, True, self.L)
        _lib.check(self.lib.nib_gp_cholesky(self.L.data_ptr(), n, self.L.stride(0), self.info.data_ptr(), st),
                   "nib_gp_cholesky")
        self.alpha_vec.copy_(self.y_d)
        for trans in (0, 1):
            _lib.check(self.lib.nib_gp_trsm(self.L.data_ptr(), n, self.L.stride(0), self.alpha_vec.data_ptr(), 1, 1,
                                            trans, st), "nib_gp_trsm")
        info = int(self.info.item())
        if info != 0:
            raise np.linalg.LinAlgError(
                f"The kernel is not returning a positive definite matrix (pivot {info}). "
                "Try gradually increasing the 'alpha' parameter of your GaussianProcessRegressor estimator.")

    # ---- sklearn surface -------------------------------------------------------------------------
    def fit(self, X, y):
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        self.kind, self.X_d, self.dim = self._upload_X(X)
        self.n = n = int(self.X_d.shape[0])
        if y.shape[0] != n:
            raise ValueError("X and y have inconsistent lengths")
        if self.normalize_y:   # _gpr.py:276-280
            self._y_train_mean = float(np.mean(y))
            std = float(np.std(y))
            self._y_train_std = std if std != 0.0 else 1.0
        else:
            self._y_train_mean, self._y_train_std = 0.0, 1.0
        yn = (y - self._y_train_mean) / self._y_train_std
        dev = self.device
        self.y_d = torch.from_numpy(yn).to(dev)
        self.L = torch.empty(n, n, dtype=torch.float64, device=dev)
        self.alpha_vec = torch.empty(n, dtype=torch.float64, device=dev)
        self.info = torch.zeros(1, dtype=torch.int32, device=dev)
        self._K0 = self._Kinv = None

        if self.optimizer is not None:
            lo, hi = np.log(self.length_scale_bounds[0]), np.log(self.length_scale_bounds[1])

            def obj(theta):
                lml, g = self.log_marginal_likelihood(theta, eval_gradient=True, _keep=False)
                return -lml, -np.atleast_1d(g)

            def run(theta0):
                r = minimize(obj, theta0, method="L-BFGS-B", jac=True, bounds=[(lo, hi)])
                return r.x, r.fun

            optima = [run(np.array([np.log(self.length_scale0)]))]
            if self.n_restarts_optimizer > 0:
                rng = np.random.RandomState(self.random_state) if not isinstance(self.random_state, np.random.RandomState) else self.random_state
                for _ in range(self.n_restarts_optimizer):
                    theta0 = rng.uniform(lo, hi, size=1)   # _gpr.py:324-327
                    optima.append(run(theta0))
            best = min(optima, key=lambda t: t[1])
            self.length_scale_ = float(np.exp(best[0][0]))
            self.log_marginal_likelihood_value_ = -float(best[1])
        else:
            self.length_scale_ = self.length_scale0
        self._factor(self.length_scale_)
        self._K0 = self._Kinv = None   # free optimiser scratch
        return self

    def log_marginal_likelihood(self, theta=None, eval_gradient: bool = False, _keep: bool = True):
        """_gpr.py:538-655 for theta = log(length_scale)."""
        ell = self.length_scale_ if theta is None else float(np.exp(np.atleast_1d(theta)[0]))
        n = self.n
        st = _lib.stream_handle()
        self.stats["lml_evals"] += 1
        try:
            self._factor(ell)
        except np.linalg.LinAlgError:
            return (-np.inf, np.zeros(1)) if eval_gradient else -np.inf
        lml = C.c_double()
        _lib.check(self.lib.nib_gp_lml(self.L.data_ptr(), n, self.L.stride(0), self.y_d.data_ptr(),
                                       self.alpha_vec.data_ptr(), C.byref(lml), st), "nib_gp_lml")
        if not eval_gradient:
            return lml.value
        dev = self.device
        if self._K0 is None:
            self._K0 = torch.empty(n, n, dtype=torch.float64, device=dev)
            self._Kinv = torch.empty(n, n, dtype=torch.float64, device=dev)
        self._gram(self.kind, self.X_d, n, self.X_d, n, self.dim, ell, 0.0, True, self._K0)
        self._Kinv.zero_()
        self._Kinv.diagonal().fill_(1.0)
        for trans in (0, 1):   # K^-1 = L^-T L^-1 I  (cho_solve(L, eye), _gpr.py:631)
            _lib.check(self.lib.nib_gp_trsm(self.L.data_ptr(), n, self.L.stride(0), self._Kinv.data_ptr(), n, n, trans, st),
                       "nib_gp_trsm")
        g = C.c_double()
        if self.kind == "bits":
            _lib.check(self.lib.nib_gp_lml_grad(self._K0.data_ptr(), self._Kinv.data_ptr(), self.alpha_vec.data_ptr(),
                                                self.X_d.data_ptr(), self.dim, n, n, ell, C.byref(g), st), "nib_gp_lml_grad")
        else:
            _lib.check(self.lib.nib_gp_lml_grad_rbf(self._K0.data_ptr(), self._Kinv.data_ptr(),
                                                    self.alpha_vec.data_ptr(), self.X_d.data_ptr(), self.dim, n, n, ell,
                                                    C.byref(g), st), "nib_gp_lml_grad_rbf")
        return lml.value, np.array([g.value])

    def predict_device(self, Xq):
        """(mu, var, std) as fp64 CUDA tensors for all queries (chunked so scratch stays bounded)."""
        kind, Q, dim = self._upload_X(Xq)
        if kind != self.kind or dim != self.dim:
            raise ValueError("query inputs do not match the training inputs")
        m = int(Q.shape[0])
        dev = self.device
        mu = torch.empty(m, dtype=torch.float64, device=dev)
        var = torch.empty(m, dtype=torch.float64, device=dev)
        sd = torch.empty(m, dtype=torch.float64, device=dev)
        n = self.n
        st = _lib.stream_handle()
        ch = min(m, self.query_chunk)
        Ks = torch.empty(ch, n, dtype=torch.float64, device=dev)
        work = torch.empty(n * ch, dtype=torch.float64, device=dev)
        for i in range(0, m, ch):
            c = min(ch, m - i)
            self._gram(kind, Q[i:i + c], c, self.X_d, n, dim, self.length_scale_, 0.0, False, Ks)
            _lib.check(self.lib.nib_gp_posterior(self.L.data_ptr(), n, self.L.stride(0), self.alpha_vec.data_ptr(),
                                                 Ks.data_ptr(), c, Ks.stride(0), self._y_train_mean, self._y_train_std,
                                                 1.0, work.data_ptr(), mu[i:i + c].data_ptr(), var[i:i + c].data_ptr(),
                                                 sd[i:i + c].data_ptr(), st), "nib_gp_posterior")
        return mu, var, sd

    def predict(self, X, return_std: bool = False):
        mu, _, sd = self.predict_device(X)
        if return_std:
            return mu.cpu().numpy(), sd.cpu().numpy()
        return mu.cpu().numpy()


def expected_improvement_device(mu: torch.Tensor, sigma: torch.Tensor, best: float, greater_is_better: bool = False):
    """+EI and its argmax on the device (BayesianOptimization.py:37-54 returns -EI)."""
    lib = _lib.load()
    m = int(mu.shape[0])
    ei = torch.empty(m, dtype=torch.float64, device=mu.device)
    arg = torch.empty(1, dtype=torch.int64, device=mu.device)
    _lib.check(lib.nib_gp_ei(mu.data_ptr(), sigma.data_ptr(), m, float(best), int(greater_is_better), ei.data_ptr(),
                             arg.data_ptr(), _lib.stream_handle()), "nib_gp_ei")
    return ei, arg


def expected_improvement(x, gaussian_process, evaluated_loss, greater_is_better=False, n_params=1):
    """Drop-in for BayesianOptimization.py:16-54 (same signature, returns -EI as numpy)."""
    x_to_predict = np.asarray(x).reshape(-1, n_params)
    mu, _, sd = gaussian_process.predict_device(x_to_predict)
    best = np.max(evaluated_loss) if greater_is_better else np.min(evaluated_loss)
    ei, _ = expected_improvement_device(mu, sd, float(best), greater_is_better)
    return -1 * ei.cpu().numpy()
