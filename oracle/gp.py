"""GP surrogate + Expected Improvement restated in numpy/scipy (test infrastructure only).

Follows scikit-learn 1.9.0 `sklearn/gaussian_process/_gpr.py` (the third-party arithmetic behind
BayesianOptimization.py:150-166; unpinned in the reference's requirements.txt, pinned here to the
installed 1.9.0) and `kernels.py:1561-1569` for the RBF.  `sklearn_gp()` builds the real estimator
exactly as BayesianOptimization.py:154-159 does so the restatement can be pinned against it.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular
from scipy.spatial.distance import cdist, pdist, squareform
from scipy.stats import norm


def bits_to_matrix(Z: np.ndarray, S: int) -> np.ndarray:
    """[N, words] uint64 selection bits -> [N, S] float64 0/1 design matrix (the GP input X)."""
    N = Z.shape[0]
    X = np.zeros((N, S), dtype=np.float64)
    for s in range(S):
        X[:, s] = ((Z[:, s // 64] >> np.uint64(s % 64)) & np.uint64(1)).astype(np.float64)
    return X


def rbf_gram(X, Y=None, length_scale=1.0):
    """kernels.py:1561-1569: exp(-0.5 * sqeuclidean(X / l)); diagonal exactly 1 when Y is None."""
    X = np.atleast_2d(X)
    if Y is None:
        d = pdist(X / length_scale, metric="sqeuclidean")
        K = squareform(np.exp(-0.5 * d))
        np.fill_diagonal(K, 1)
        return K
    return np.exp(-0.5 * cdist(X / length_scale, Y / length_scale, metric="sqeuclidean"))


def gp_fit(X, y, length_scale, alpha=1e-5, normalize_y=True):
    """_gpr.py:276-280 (normalise), :349-367 (K + alpha I, cholesky lower, cho_solve)."""
    y = np.asarray(y, dtype=np.float64)
    if normalize_y:
        y_mean = np.mean(y, axis=0)
        y_std = np.std(y, axis=0)
        if y_std == 0.0:  # _handle_zeros_in_scale
            y_std = 1.0
        yn = (y - y_mean) / y_std
    else:
        y_mean, y_std, yn = 0.0, 1.0, y
    K = rbf_gram(X, None, length_scale)
    K[np.diag_indices_from(K)] += alpha
    L = cholesky(K, lower=True, check_finite=False)
    a = cho_solve((L, True), yn, check_finite=False)
    return dict(X=np.asarray(X, dtype=np.float64), L=L, alpha=a, y_mean=float(y_mean), y_std=float(y_std),
                yn=yn, length_scale=float(length_scale))


def gp_predict(fit, Xq):
    """_gpr.py:446-496: mean, and std via V = solve_triangular(L, K_trans.T); 1 - sum(V^2), clip."""
    Ks = rbf_gram(Xq, fit["X"], fit["length_scale"])
    mu = fit["y_std"] * (Ks @ fit["alpha"]) + fit["y_mean"]
    V = solve_triangular(fit["L"], Ks.T, lower=True, check_finite=False)
    var = np.ones(Ks.shape[0]) - np.einsum("ij,ji->i", V.T, V)
    var[var < 0] = 0.0
    var = var * fit["y_std"] ** 2
    return mu, var, np.sqrt(var)


def lml_and_grad(X, yn, length_scale, alpha=1e-5):
    """_gpr.py:588-655 with theta = log(length_scale): returns (lml, dlml/dtheta)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    K0 = rbf_gram(X, None, length_scale)
    K = K0.copy()
    K[np.diag_indices_from(K)] += alpha
    L = cholesky(K, lower=True, check_finite=False)
    a = cho_solve((L, True), yn, check_finite=False)
    lml = -0.5 * yn @ a - np.log(np.diag(L)).sum() - n / 2 * np.log(2 * np.pi)
    D2 = squareform(pdist(X / length_scale, metric="sqeuclidean"))
    dK = K0 * D2  # kernels.py:1571-1576
    Kinv = cho_solve((L, True), np.eye(n), check_finite=False)
    grad = 0.5 * np.einsum("ij,ji->", np.outer(a, a) - Kinv, dK)
    return float(lml), float(grad)


def expected_improvement(mu, sigma, evaluated_loss, greater_is_better=False):
    """BayesianOptimization.py:37-54 — returns -EI like the reference (line :52 is a no-op `==`)."""
    loss_optimum = np.max(evaluated_loss) if greater_is_better else np.min(evaluated_loss)
    scaling_factor = (-1) ** (not greater_is_better)
    with np.errstate(divide="ignore", invalid="ignore"):
        Z = scaling_factor * (mu - loss_optimum) / sigma
        ei = scaling_factor * (mu - loss_optimum) * norm.cdf(Z) + sigma * norm.pdf(Z)
    return -1 * ei


def sklearn_gp(alpha=1e-5, n_restarts_optimizer=10, optimizer="fmin_l_bfgs_b", random_state=None, length_scale=1.0):
    """The estimator of BayesianOptimization.py:154-159."""
    import sklearn.gaussian_process as gp

    kernel = gp.kernels.RBF(length_scale=length_scale)
    return gp.GaussianProcessRegressor(kernel=kernel, alpha=alpha, n_restarts_optimizer=n_restarts_optimizer,
                                       normalize_y=True, optimizer=optimizer, random_state=random_state)
