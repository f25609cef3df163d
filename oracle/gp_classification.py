"""Variational GP classification of gp_classification.py, restated in numpy.  Test infrastructure only.

The reference builds `GPClassificationModel(gpytorch.models.GridInducingVariationalGP)` with grid_size = 10 over [0, n]^2,
ConstantMean(bounds +-1e-5), RBFKernel scaled by exp(log_outputscale) (gp_classification.py:139-156), a BernoulliLikelihood
(:317) and runs 30 Adam(lr = 0.1) steps on -VariationalMarginalLogLikelihood (:160-217); prediction is
`likelihood(model(x)).mean()` over the n x n pixels (:241-253).  gpytorch is absent from this image and unpinned by the
reference (pre-0.1 API: `gpytorch.random_variables`, `GridInducingVariationalGP`): **parity unpinned**.  This file states
the model by its dense DEFINITION (dense interpolation matrix W, dense covariances, scipy special functions), an
independent formulation of what the product computes with sparse stencils and quadrature kernels (csrc/ski.cu):

    u ~ N(0, K_UU) on the grid, q(u) = N(m, S), S = Ls Ls^T          f = c + W u          p(y | f) = Phi(y f)
    ELBO = sum_i E_{N(mu_i, s2_i)}[log Phi(y_i f)] - KL(q(u) || p(u)),   loss = -ELBO / n   (VariationalMarginalLogLikelihood
    divides both terms by n_data).  Expectations: Gauss-Hermite quadrature (the reference's likelihood SAMPLES f, which
    no test could pin); derivatives by Bonnet / Price: dE/dmu = E[d/df], dE/ds2 = 1/2 E[d^2/df^2].
"""
from __future__ import annotations

import numpy as np
from scipy.special import log_ndtr, ndtr

from .ski import grid_kernel, interp_matrix, make_grid


def expected_loglik(mu, s2, y, nodes=64):
    """E[log Phi(y f)], f ~ N(mu, s2), and its derivatives w.r.t. mu and s2."""
    x, w = np.polynomial.hermite.hermgauss(nodes)
    f = mu[:, None] + np.sqrt(2.0 * s2)[:, None] * x[None, :]
    z = y[:, None] * f
    lp = log_ndtr(z)
    r = np.exp(-0.5 * z * z - lp) / np.sqrt(2 * np.pi)            # phi / Phi
    wn = w[None, :] / np.sqrt(np.pi)
    e = (wn * lp).sum(1)
    dmu = (wn * (y[:, None] * r)).sum(1)
    ds2 = (wn * (-0.5 * y[:, None] ** 2 * r * (z + r))).sum(1)
    return e, dmu, ds2


def elbo_terms(X, y, m, Ls, bounds=(0.0, 224.0), grid_size=10, length_scale=1.0, outputscale=1.0, const_mean=0.0, jitter=1e-6,
               nodes=64):
    """(sum of expected log-likelihoods, KL, grad_m of the data term, grad_S of the data term)."""
    g0, h = make_grid(bounds[0], bounds[1], grid_size)
    K = grid_kernel(g0, h, grid_size, length_scale, outputscale)
    K[np.diag_indices_from(K)] += jitter * outputscale
    W = interp_matrix(X, g0, h, grid_size)
    S = Ls @ Ls.T
    mu = const_mean + W @ m
    s2 = np.einsum("ij,jk,ik->i", W, S, W)
    e, dmu, ds2 = expected_loglik(mu, np.maximum(s2, 0.0), np.asarray(y, dtype=np.float64), nodes=nodes)
    G = K.shape[0]
    Kinv = np.linalg.inv(K)
    kl = 0.5 * (np.trace(Kinv @ S) + m @ Kinv @ m - G + np.linalg.slogdet(K)[1] - np.linalg.slogdet(S)[1])
    return e.sum(), kl, W.T @ dmu, np.einsum("i,ij,ik->jk", ds2, W, W)


def predict_prob(Xq, m, Ls, bounds=(0.0, 224.0), grid_size=10, const_mean=0.0):
    """E_q[Phi(f)] = Phi(mu / sqrt(1 + s2)) at the query points."""
    g0, h = make_grid(bounds[0], bounds[1], grid_size)
    W = interp_matrix(Xq, g0, h, grid_size)
    S = Ls @ Ls.T
    mu = const_mean + W @ m
    s2 = np.maximum(np.einsum("ij,jk,ik->i", W, S, W), 0.0)
    return ndtr(mu / np.sqrt(1.0 + s2)), mu, s2
