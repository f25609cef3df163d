"""CPU: the dense-definition oracle of the variational GP classifier is self-consistent - its analytic gradients of the
data term (Bonnet / Price derivatives through the quadrature) equal central differences of its own ELBO, and the closed
form E[Phi(f)] equals quadrature.  (gpytorch is absent: parity of this stage is unpinned, see oracle/gp_classification.py.)"""
import numpy as np
from scipy.special import ndtr

from oracle import gp_classification as ovgp


def test_data_term_gradients_match_finite_differences():
    rng = np.random.RandomState(0)
    n, gs = 60, 6
    G = gs * gs
    X = rng.rand(n, 2) * 50.0
    y = rng.randint(-2, 4, size=n).astype(np.float64)
    m = rng.randn(G) * 0.2
    Ls = np.tril(rng.randn(G, G) * 0.05) + np.eye(G) * 0.7
    e, kl, gm, gS = ovgp.elbo_terms(X, y, m, Ls, (0.0, 50.0), gs, 12.0, 1.3)
    h = 1e-6
    for k in (0, 7, 20, 35):
        d = np.zeros(G); d[k] = h
        fd = (ovgp.elbo_terms(X, y, m + d, Ls, (0.0, 50.0), gs, 12.0, 1.3)[0] - ovgp.elbo_terms(X, y, m - d, Ls, (0.0, 50.0), gs, 12.0, 1.3)[0]) / (2 * h)
        np.testing.assert_allclose(gm[k], fd, rtol=1e-5, atol=1e-7)
    # dE/dLs = 2 * sym(gS) Ls
    gL = np.tril((gS + gS.T) @ Ls)
    for (a, b) in ((3, 3), (10, 2), (30, 30), (35, 0)):
        d = np.zeros((G, G)); d[a, b] = h
        fd = (ovgp.elbo_terms(X, y, m, Ls + d, (0.0, 50.0), gs, 12.0, 1.3)[0] - ovgp.elbo_terms(X, y, m, Ls - d, (0.0, 50.0), gs, 12.0, 1.3)[0]) / (2 * h)
        np.testing.assert_allclose(gL[a, b], fd, rtol=1e-4, atol=1e-7)


def test_predictive_probability_closed_form_equals_quadrature():
    rng = np.random.RandomState(1)
    mu, s2 = rng.randn(50), rng.rand(50) * 3.0
    x, w = np.polynomial.hermite.hermgauss(80)
    quad = (w[None, :] / np.sqrt(np.pi) * ndtr(mu[:, None] + np.sqrt(2 * s2)[:, None] * x[None, :])).sum(1)
    np.testing.assert_allclose(ndtr(mu / np.sqrt(1 + s2)), quad, rtol=1e-10)
