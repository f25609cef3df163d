"""GPU: stage 3 (GP surrogate + EI) through the C ABI, against the scikit-learn 1.9.0 fixture, the reference's
own `expected_improvement` outputs, and the numpy oracle.  Tolerance: north_star's 1e-4 on posterior mean and
variance (written below); the fp64 kernels land many orders of magnitude inside it."""
import os

import numpy as np
import pytest
import torch

from oracle import gp as ogp
from oracle import masks as om

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=True)


def test_fixed_theta_fit_predict_vs_sklearn_fixture(nib, golden_dir):
    g = _g(golden_dir, "gp_sklearn.npz")
    ell = float(g["length_scale"])
    gp = nib.GaussianProcessRegressor(alpha=1e-5, normalize_y=True, length_scale=ell, optimizer=None)
    gp.fit(g["Xt"], g["yt"])
    n = g["Xt"].shape[0]
    L = torch.tril(gp.L).cpu().numpy()
    np.testing.assert_allclose(L, g["L"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(gp.alpha_vec.cpu().numpy(), g["alpha_vec"], rtol=1e-5, atol=1e-7)
    mu, sd = gp.predict(g["Xq"], return_std=True)
    assert np.abs(mu - g["mu"]).max() <= TOL and np.abs(sd ** 2 - g["std"] ** 2).max() <= TOL
    np.testing.assert_allclose(mu, g["mu"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(sd, g["std"], rtol=1e-5, atol=1e-8)
    for t, l, dl in zip(g["thetas"], g["lml"], g["lml_grad"]):
        lml, grad = gp.log_marginal_likelihood(np.array([t]), eval_gradient=True)
        np.testing.assert_allclose(lml, l, rtol=1e-8)
        np.testing.assert_allclose(grad[0], dl, rtol=1e-6, atol=1e-8)
    assert n == gp.n


def test_hyperparameter_search_reaches_sklearn_optimum(nib, golden_dir):
    g = _g(golden_dir, "gp_sklearn.npz")
    gp = nib.GaussianProcessRegressor(alpha=1e-5, normalize_y=True, n_restarts_optimizer=10, random_state=0)
    gp.fit(g["Xt"], g["yt"])
    assert gp.log_marginal_likelihood_value_ >= float(g["lml_opt"]) - 1e-5
    np.testing.assert_allclose(gp.length_scale_, float(g["length_scale"]), rtol=1e-3)
    mu, sd = gp.predict(g["Xq"], return_std=True)
    assert np.abs(mu - g["mu"]).max() <= TOL and np.abs(sd ** 2 - g["std"] ** 2).max() <= TOL


def test_reference_1d_first_index_gp(nib, golden_dir):
    """The GP exactly as the reference uses it: 13 scalar firstIndex inputs (BayesianOptimization.py:137-166)."""
    g = _g(golden_dir, "gp_sklearn_1d.npz")
    gp = nib.GaussianProcessRegressor(alpha=1e-5, normalize_y=True, length_scale=float(g["length_scale"]), optimizer=None)
    gp.fit(g["x"], g["y"])
    mu, sd = gp.predict(g["xq"], return_std=True)
    assert np.abs(mu - g["mu"]).max() <= TOL and np.abs(sd ** 2 - g["std"] ** 2).max() <= TOL
    gp2 = nib.GaussianProcessRegressor(alpha=1e-5, normalize_y=True, n_restarts_optimizer=10, random_state=1).fit(g["x"], g["y"])
    assert gp2.log_marginal_likelihood_value_ >= float(g["lml_opt"]) - 1e-5


def test_expected_improvement_vs_reference_outputs(nib, golden_dir):
    g = _g(golden_dir, "ei.npz")
    mu = torch.from_numpy(g["mu"]).cuda()
    sg = torch.from_numpy(g["sigma"]).cuda()
    for gib, key in ((True, "neg_ei_max"), (False, "neg_ei_min")):
        best = g["losses"].max() if gib else g["losses"].min()
        ei, arg = nib.expected_improvement_device(mu, sg, float(best), gib)
        got = -ei.cpu().numpy()
        want = g[key]
        assert np.array_equal(np.isnan(got), np.isnan(want))   # sigma == 0 -> NaN, exactly like the reference
        ok = ~np.isnan(want)
        np.testing.assert_allclose(got[ok], want[ok], rtol=1e-12, atol=1e-15)
        assert int(arg.item()) == int(np.nanargmax(-want))


@pytest.mark.parametrize("n,m,S", [(1, 3, 50), (63, 65, 50), (1000, 513, 50), (300, 40, 130)])
def test_sizes_vs_oracle(nib, n, m, S):
    rng = np.random.RandomState(n + m)
    sels = [list(rng.choice(S - 1, size=max(1, int(0.4 * S)), replace=False)) for _ in range(n + m)]
    Z = om.selection_bits(sels, S)
    X = ogp.bits_to_matrix(Z, S)
    y = rng.rand(n)
    ell = 3.0
    fit = ogp.gp_fit(X[:n], y, ell)
    mu0, var0, sd0 = ogp.gp_predict(fit, X[n:])
    gp = nib.GaussianProcessRegressor(alpha=1e-5, length_scale=ell, optimizer=None, query_chunk=256)
    gp.fit(Z[:n], y)
    mu, var, sd = (t.cpu().numpy() for t in gp.predict_device(Z[n:]))
    assert np.abs(mu - mu0).max() <= TOL and np.abs(var - var0).max() <= TOL
    np.testing.assert_allclose(mu, mu0, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(var, var0, rtol=1e-4, atol=1e-8)
    if n > 1:
        lml0, g0 = ogp.lml_and_grad(X[:n], fit["yn"], ell)
        lml, g1 = gp.log_marginal_likelihood(np.log([ell]), eval_gradient=True)
        np.testing.assert_allclose(lml, lml0, rtol=1e-8)
        np.testing.assert_allclose(g1[0], g0, rtol=1e-6, atol=1e-8)


def test_n8192_vs_oracle(nib):
    """BASELINE configs[3] size, n = m = 8192: the 512-wide Cholesky super-blocks and the recursive TRSM (which only engage
    at large n) against the scikit-learn-pinned oracle (scipy cholesky / solve_triangular on the host, ~20 s)."""
    n = m = 8192
    S, ell = 50, 3.0
    rng = np.random.RandomState(8192)
    sels = [list(rng.choice(S - 1, size=20, replace=False)) for _ in range(n + m)]
    Z = om.selection_bits(sels, S)
    X = ogp.bits_to_matrix(Z, S)
    y = rng.rand(n)
    fit = ogp.gp_fit(X[:n], y, ell)
    mu0, var0, sd0 = ogp.gp_predict(fit, X[n:])
    gp = nib.GaussianProcessRegressor(alpha=1e-5, length_scale=ell, optimizer=None, query_chunk=8192)
    gp.fit(Z[:n], y)
    mu, var, sd = (t.cpu().numpy() for t in gp.predict_device(Z[n:]))
    assert np.abs(mu - mu0).max() <= TOL, np.abs(mu - mu0).max()
    assert np.abs(var - var0).max() <= TOL, np.abs(var - var0).max()
    L = torch.tril(gp.L).cpu().numpy()
    assert np.abs(L - fit["L"]).max() <= 1e-7, np.abs(L - fit["L"]).max()
    np.testing.assert_allclose(gp.alpha_vec.cpu().numpy(), fit["alpha"], rtol=1e-5, atol=1e-5 * np.abs(fit["alpha"]).max())
    ei, arg = nib.expected_improvement_device(torch.from_numpy(mu).cuda(), torch.from_numpy(sd).cuda(), float(y.max()), True)
    ref_ei = -ogp.expected_improvement(mu0, sd0, y, greater_is_better=True)
    assert int(arg.item()) == int(np.nanargmax(ref_ei))


def test_log_marginal_likelihood_leaves_the_fitted_model_untouched(nib):
    """scikit-learn's log_marginal_likelihood(theta) is side-effect free (_gpr.py:538-655): querying another theta on a
    fitted model must not change what predict() returns (it used to overwrite L and alpha in place)."""
    S, n, m = 50, 300, 64
    rng = np.random.RandomState(5)
    sels = [list(rng.choice(S - 1, size=20, replace=False)) for _ in range(n + m)]
    Z = om.selection_bits(sels, S)
    y = rng.rand(n)
    gp = nib.GaussianProcessRegressor(alpha=1e-5, length_scale=3.0, optimizer=None).fit(Z[:n], y)
    mu_a, sd_a = gp.predict(Z[n:], return_std=True)
    lml_fit = gp.log_marginal_likelihood()
    lml_other, g = gp.log_marginal_likelihood(np.log([0.7]), eval_gradient=True)
    assert lml_other != lml_fit
    mu_b, sd_b = gp.predict(Z[n:], return_std=True)
    assert np.array_equal(mu_a, mu_b) and np.array_equal(sd_a, sd_b)
    assert gp.log_marginal_likelihood() == lml_fit
    lml0, g0 = ogp.lml_and_grad(ogp.bits_to_matrix(Z[:n], S), (y - y.mean()) / y.std(), 0.7)
    np.testing.assert_allclose(lml_other, lml0, rtol=1e-8)
    np.testing.assert_allclose(g[0], g0, rtol=1e-6, atol=1e-8)


def test_duplicate_masks_without_jitter_raise_like_sklearn(nib):
    Z = om.selection_bits([[1, 2, 3], [1, 2, 3], [4, 5]], 50)
    gp = nib.GaussianProcessRegressor(alpha=0.0, length_scale=1.0, optimizer=None)
    with pytest.raises(np.linalg.LinAlgError):
        gp.fit(Z, np.array([0.1, 0.2, 0.3]))


def test_cholesky_round_trip_large(nib):
    """n = 2048: L L^T reproduces K + alpha I (size-independent property); solves are consistent."""
    rng = np.random.RandomState(0)
    S, n = 50, 2048
    sels = [list(rng.choice(S - 1, size=20, replace=False)) for _ in range(n)]
    Z = om.selection_bits(sels, S)
    y = rng.rand(n)
    gp = nib.GaussianProcessRegressor(alpha=1e-5, length_scale=2.0, optimizer=None).fit(Z, y)
    L = torch.tril(gp.L)
    K = L @ L.t()
    X = torch.from_numpy(ogp.bits_to_matrix(Z, S)).cuda()
    d2 = torch.cdist(X, X) ** 2
    K0 = torch.exp(-0.5 * d2 / 4.0) + 1e-5 * torch.eye(n, dtype=torch.float64, device="cuda")
    assert (K - K0).abs().max().item() < 1e-10
    r = K0 @ gp.alpha_vec - gp.y_d
    assert r.abs().max().item() < 1e-6
    # training points are interpolated up to the jitter
    mu, sd = gp.predict(Z[:64], return_std=True)
    assert np.abs(mu - y[:64]).max() < 1e-2 and sd.max() < 0.05


@pytest.mark.parametrize("n0,m,rounds", [(5, 40, 12), (200, 300, 70), (1000, 513, 5)])
def test_rank_one_acquisition_updates_equal_refit_and_oracle(nib, n0, m, rounds):
    """ActiveMaskGP: `rounds` appended candidates (crossing 64-row block boundaries) == a from-scratch fit on the grown
    training set, both ours and the scikit-learn-pinned oracle; used candidates are masked out of the pool."""
    S, ell = 50, 3.0
    rng = np.random.RandomState(n0 + m)
    sels, seen = [], set()
    while len(sels) < n0 + m:                       # distinct masks (a duplicate makes K singular up to alpha)
        s = tuple(sorted(rng.choice(S - 1, size=20, replace=False)))
        if s not in seen:
            seen.add(s)
            sels.append(list(s))
    Z = om.selection_bits(sels, S)
    y0 = rng.rand(n0)
    agp = nib.ActiveMaskGP(Z[n0:], alpha=1e-5, length_scale=ell, capacity=rounds).fit(Z[:n0], y0)
    picks = [int(p) for p in rng.choice(m, size=rounds, replace=False)]
    ynew = rng.rand(rounds)
    for p, yv in zip(picks, ynew):
        agp.append(p, float(yv))
    mu, var, sd = (t.cpu().numpy() for t in agp.posterior())
    Zt = np.concatenate([Z[:n0], Z[n0:][picks]], 0)
    yt = np.concatenate([y0, ynew])
    fit = ogp.gp_fit(ogp.bits_to_matrix(Zt, S), yt, ell)
    mu0, var0, sd0 = ogp.gp_predict(fit, ogp.bits_to_matrix(Z[n0:], S))
    live = np.ones(m, bool)
    live[picks] = False
    assert np.isnan(sd[~live]).all() and not np.isnan(sd[live]).any()
    assert np.abs(mu - mu0).max() <= TOL and np.abs(var - var0).max() <= TOL
    np.testing.assert_allclose(mu, mu0, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(sd[live], sd0[live], rtol=1e-4, atol=1e-6)
    ref = nib.GaussianProcessRegressor(alpha=1e-5, length_scale=ell, optimizer=None).fit(Zt, yt)
    mu1, var1, _ = (t.cpu().numpy() for t in ref.predict_device(Z[n0:]))
    np.testing.assert_allclose(mu, mu1, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(var, var1, rtol=1e-5, atol=1e-9)
    with pytest.raises(RuntimeError):
        agp.append(int(np.nonzero(live)[0][0]), 0.5)   # capacity exhausted


def test_bo_loop_rank_one_path_picks_the_same_candidates_as_refitting(nib):
    import BayesianOptimization as bo
    S, n0, m = 50, 150, 200
    rng = np.random.RandomState(3)
    sels = [list(rng.choice(S - 1, size=20, replace=False)) for _ in range(n0 + m)]
    Z = om.selection_bits(sels, S)
    w = rng.rand(S)
    score = lambda b: np.array([float(sum(w[s] for s in range(S) if (int(v[0]) >> s) & 1)) / 10.0 for v in b])
    y0 = score(Z[:n0])
    _, ya, ha = bo.bayesian_optimisation_masks(8, score, Z[:n0], y0, Z[n0:], length_scale=3.0)
    _, yb, hb = bo.bayesian_optimisation_masks(8, score, Z[:n0], y0, Z[n0:], length_scale=3.0, refit_every=10 ** 9)
    assert [h["candidate"] for h in ha] == [h["candidate"] for h in hb]
    np.testing.assert_allclose(ya, yb)
    np.testing.assert_allclose([h["ei"] for h in ha], [h["ei"] for h in hb], rtol=1e-5, atol=1e-9)
