"""GPU: variational GP classification on the inducing grid (gp_classification.py) against the dense definition in
oracle/gp_classification.py.  Parity unpinned (gpytorch absent, pre-0.1 API): the oracle is an independent formulation
(dense interpolation matrix, dense covariances, scipy special functions, 64-node quadrature) of the same model."""
import numpy as np
import pytest
import torch

from oracle import gp_classification as ovgp

pytestmark = pytest.mark.gpu


def _problem(n, gs, seed, size=224.0):
    rng = np.random.RandomState(seed)
    X = rng.randint(0, int(size), size=(n, 2)).astype(np.float64)
    y = rng.randint(0, 6, size=n).astype(np.float64)            # summed mask labels (the reference feeds counts, :78-80)
    y[rng.rand(n) < 0.3] *= -1.0                                # and +-1-style labels: any real y must work
    G = gs * gs
    m = rng.randn(G) * 0.3
    Ls = np.tril(rng.randn(G, G) * 0.05) + np.eye(G) * 0.8
    return X, y, m, Ls


@pytest.mark.parametrize("n,gs", [(300, 10), (2000, 10), (500, 14)])
def test_expected_loglik_and_gradients_equal_dense_definition(nib, n, gs):
    from network_interpretation_imagenet_b200.vgp import GridVariationalGPClassifier
    X, y, m, Ls = _problem(n, gs, n + gs)
    clf = GridVariationalGPClassifier(gs, ((0.0, 224.0), (0.0, 224.0)), log_lengthscale=np.log(40.0), log_outputscale=0.3,
                                      const_mean=0.0)
    clf.variational_mean, clf.chol_variational_covar = m.copy(), Ls.copy()
    ell, gm, gS, gc = clf.data_term(clf._dev64(X), clf._dev64(y))
    # same 20-node rule, independent (dense, scipy) code: tight; the converged 64-node rule: what 20 nodes leave on the table
    e0, kl0, gm0, gS0 = ovgp.elbo_terms(X, y, m, Ls, (0.0, 224.0), gs, 40.0, np.exp(0.3), 0.0, jitter=clf.jitter, nodes=20)
    assert abs(ell - e0) <= 1e-10 * max(1.0, abs(e0))
    np.testing.assert_allclose(gm, gm0, rtol=1e-8, atol=1e-10 * np.abs(gm0).max())
    np.testing.assert_allclose(gS + gS.T, gS0 + gS0.T, rtol=1e-8, atol=1e-10 * np.abs(gS0).max())
    e1, _, gm1, _ = ovgp.elbo_terms(X, y, m, Ls, (0.0, 224.0), gs, 40.0, np.exp(0.3), 0.0, jitter=clf.jitter, nodes=64)
    assert abs(ell - e1) <= 1e-4 * max(1.0, abs(e1))
    np.testing.assert_allclose(gm, gm1, rtol=0, atol=1e-2 * np.abs(gm1).max())
    kl, g_m, g_Ls, g_logl, g_logos = clf.kl_term()
    assert abs(kl - kl0) <= 1e-8 * max(1.0, abs(kl0))
    # hyper-parameter gradients of the KL by central differences on the dense definition
    for name, g in (("log_lengthscale", g_logl), ("log_outputscale", g_logos)):
        h = 1e-5
        vals = []
        for s in (+1, -1):
            ll = np.log(40.0) + (s * h if name == "log_lengthscale" else 0.0)
            lo = 0.3 + (s * h if name == "log_outputscale" else 0.0)
            vals.append(ovgp.elbo_terms(X[:8], y[:8], m, Ls, (0.0, 224.0), gs, np.exp(ll), np.exp(lo), 0.0, jitter=clf.jitter)[1])
        # (central differences of a log-determinant of an ill-conditioned K: G = 196 grid points at length scale 40)
        np.testing.assert_allclose(g, (vals[0] - vals[1]) / (2 * h), rtol=2e-3, atol=1e-6)


def test_predictive_probability_equals_dense_definition(nib):
    from network_interpretation_imagenet_b200.vgp import GridVariationalGPClassifier
    X, y, m, Ls = _problem(10, 10, 3)
    clf = GridVariationalGPClassifier(10, ((0.0, 224.0), (0.0, 224.0)))
    clf.variational_mean, clf.chol_variational_covar = m.copy(), Ls.copy()
    Xq = np.stack(np.meshgrid(np.arange(0, 224, 5.0), np.arange(0, 224, 3.0), indexing="ij"), -1).reshape(-1, 2)
    p, mu, var = (t.cpu().numpy() for t in clf.predict_proba(Xq, return_latent=True))
    p0, mu0, var0 = ovgp.predict_prob(Xq, m, Ls, (0.0, 224.0), 10)
    assert np.abs(p - p0).max() <= 1e-10 and np.abs(mu - mu0).max() <= 1e-10 and np.abs(var - var0).max() <= 1e-10


def test_thirty_adam_steps_lower_the_loss_and_separate_the_classes(nib):
    """The reference's training schedule (30 Adam steps, lr 0.1) on a synthetic heat map: pixels in a disc carry label +1,
    the rest -1.  The loss must fall and the predictive probability must separate the two regions."""
    from network_interpretation_imagenet_b200.vgp import GridVariationalGPClassifier
    ii, jj = np.meshgrid(np.arange(0, 224, 2.0), np.arange(0, 224, 2.0), indexing="ij")
    X = np.stack([ii, jj], -1).reshape(-1, 2)
    inside = ((ii - 110) ** 2 + (jj - 90) ** 2 < 60 ** 2).reshape(-1)
    y = np.where(inside, 1.0, -1.0)
    clf = GridVariationalGPClassifier(10, ((0.0, 224.0), (0.0, 224.0)), log_lengthscale=np.log(30.0))
    clf.fit(X, y, num_training_iterations=30, lr=0.1, verbose=False)
    losses = [h["loss"] for h in clf.history]
    assert losses[-1] < 0.6 * losses[0], losses
    p = clf.predict_proba(X).cpu().numpy()
    assert p[inside].mean() > 0.8 and p[~inside].mean() < 0.2


def test_gp_classification_entry_point_round_trip(nib, tmp_path, monkeypatch):
    """The drop-in gp_classification.py on a ./masks directory in the reference's format: Train writes the checkpoint the
    reference's path names, Eval reloads it and returns n*n probabilities."""
    import cv2
    import importlib
    monkeypatch.chdir(tmp_path)
    (tmp_path / "masks").mkdir()
    rng = np.random.RandomState(1)
    gpc = importlib.import_module("gp_classification")
    monkeypatch.setattr(gpc, "n", 28)
    for i in range(40):
        mask = np.zeros((28, 28), np.uint8)
        r0, c0 = rng.randint(0, 14, size=2)
        mask[r0:r0 + 14, c0:c0 + 14] = 255
        label = int(r0 < 7 and c0 < 7)
        cv2.imwrite(str(tmp_path / "masks" / f"mask_{i}_{label}.png"), mask)
    model, lik = gpc.GPClassificationModel(), gpc.BernoulliLikelihood()
    # GPClassificationModel reads the module-level n at construction, like the reference
    assert model.gs == 10
    model = gpc.GridVariationalGPClassifier(10, ((0.0, 28.0), (0.0, 28.0)))
    tx, ty = gpc.prepare_training_data()
    assert tx.shape[1] == 2 and tx.shape[0] == ty.shape[0] > 0
    gpc.train(tx, ty, model, lik)
    assert (tmp_path / "gp_saved_checkpoints" / "imgenet100_epoch10_gp_cls_checkpoint.pth.tar").exists()
    model2 = gpc.GridVariationalGPClassifier(10, ((0.0, 28.0), (0.0, 28.0)))
    pred = gpc.eval_superpixels(model2, lik)
    assert pred.shape == (28 * 28,) and np.isfinite(pred).all() and (pred >= 0).all() and (pred <= 1).all()
    gpc.plot_result(pred)
    assert (tmp_path / "weighted_mask" / "predicted_class_probability_heatmap.png").exists()
