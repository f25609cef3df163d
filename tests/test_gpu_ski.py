"""GPU: pixel-coordinate GP regression on an inducing grid (gp_regression.py) and the PNG heat map, against the dense
definition in oracle/ski.py.  Parity unpinned (gpytorch absent): the oracle is an independent formulation of the same model."""
import numpy as np
import pytest
import torch

from oracle import ski as oski

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,gs,ell,noise,os_", [(300, 12, 1.0, 1.0, 1.0), (500, 30, 1.0, 1.0, 1.0), (400, 16, 25.0, 0.1, 2.0)])
def test_grid_gp_regression_equals_dense_ski_definition(nib, n, gs, ell, noise, os_):
    from network_interpretation_imagenet_b200.ski import GridGPRegression
    rng = np.random.RandomState(n)
    X = rng.randint(0, 224, size=(n, 2)).astype(np.float64)          # pixel coordinates, duplicates allowed
    y = rng.rand(n) * 40.0
    Xq = np.stack(np.meshgrid(np.arange(0, 224, 9.0), np.arange(0, 224, 7.0), indexing="ij"), -1).reshape(-1, 2)
    mean0, var0 = oski.ski_posterior(X, y, Xq, (0.0, 224.0), gs, ell, os_, noise, const_mean=0.5)
    gp = GridGPRegression(gs, ((0.0, 224.0), (0.0, 224.0)), ell, os_, noise, const_mean=0.5).fit(X, y)
    mean, var = (t.cpu().numpy() for t in gp.predict(Xq))
    scale = max(1.0, np.abs(mean0).max())
    assert np.abs(mean - mean0).max() <= 1e-4 * scale           # north-star GP tolerance
    assert np.abs(var - var0).max() <= 1e-4 * max(1.0, np.abs(var0).max())
    lat = gp.predict(Xq, likelihood=False)[1].cpu().numpy()
    np.testing.assert_allclose(lat + noise, var, rtol=1e-12)


def test_full_image_fit_and_predict_runs_at_reference_size(nib):
    """n = 224^2 training pixels, all 224^2 queries, grid 30 (the reference's configuration): finite, smooth, bounded by
    the data range, and the variance is smallest where there is data."""
    from network_interpretation_imagenet_b200.ski import GridGPRegression
    ii, jj = np.meshgrid(np.arange(224.0), np.arange(224.0), indexing="ij")
    X = np.stack([ii, jj], -1).reshape(-1, 2)
    y = (20.0 * np.exp(-((ii - 100) ** 2 + (jj - 120) ** 2) / (2 * 40.0 ** 2))).reshape(-1)
    keep = (jj.reshape(-1) < 180)                                      # no training pixels in the right-hand strip
    gp = GridGPRegression(30, ((0.0, 224.0), (0.0, 224.0)), 10.0, 50.0, 1.0).fit(X[keep], y[keep])
    mean, var = (t.cpu().numpy().reshape(224, 224) for t in gp.predict(X, likelihood=False))
    assert np.isfinite(mean).all() and np.isfinite(var).all() and (var >= -1e-9).all()
    assert np.abs(mean[:, :170] - y.reshape(224, 224)[:, :170]).max() < 1.0
    assert var[:, 200:].mean() > 10 * var[:, :170].mean()


def test_heatmap_from_png_masks(nib):
    from network_interpretation_imagenet_b200.ski import heatmap_from_masks
    rng = np.random.RandomState(0)
    masks = (rng.rand(257, 28, 36) > 0.6).astype(np.uint8) * 255
    labels = rng.randint(0, 2, size=257)
    want = oski.heatmap_from_masks(masks, labels)
    got = heatmap_from_masks(masks, labels).cpu().numpy()
    assert np.array_equal(got, want.astype(np.float32))             # integer sums: exact


def test_gp_regression_entry_point_round_trip(nib, tmp_path, monkeypatch):
    """The drop-in gp_regression.py on a ./masks directory written in the reference's format (mask_{i}_{label}.png)."""
    import cv2
    import importlib
    monkeypatch.chdir(tmp_path)
    (tmp_path / "masks").mkdir()
    rng = np.random.RandomState(1)
    masks, labels = [], []
    for i in range(40):
        m = np.zeros((224, 224), np.uint8)
        y0, x0 = rng.randint(0, 150, size=2)
        m[y0:y0 + 74, x0:x0 + 74] = 255
        lab = int(rng.rand() < 0.5)
        cv2.imwrite(str(tmp_path / "masks" / f"mask_{i}_{lab}.png"), m)
        masks.append(m)
        labels.append(lab)
    import gp_regression as gpr
    gpr = importlib.reload(gpr)
    tx, ty = gpr.prepare_training_data()
    heat = oski.heatmap_from_masks(np.stack(masks), np.asarray(labels))
    cov = (np.stack(masks) == 255).any(0)
    assert tx.shape == (int(cov.sum()), 2) and np.array_equal(ty.cpu().numpy(), heat[cov].astype(np.float32))
    assert np.array_equal(tx.cpu().numpy(), np.argwhere(cov).astype(np.float32))
    lik = gpr.GaussianLikelihood().cuda()
    model = gpr.GPRegressionModel(tx, ty, lik).cuda()
    gpr.train(tx, ty, model)
    pred = gpr.eval_superpixels(model, lik)
    assert pred.shape == (224 * 224,) and np.isfinite(pred).all()
    sub = rng.choice(tx.shape[0], size=600, replace=False)                   # dense check on a subsample of the pixels
    Xs, ys = tx.cpu().numpy()[sub].astype(np.float64), ty.cpu().numpy()[sub].astype(np.float64)
    small = gpr.GPRegressionModel(torch.from_numpy(Xs), torch.from_numpy(ys), lik)
    q = np.array([[10.0, 10.0], [100.0, 120.0], [223.0, 0.0], [57.0, 199.0]])
    mean, var = (t.cpu().numpy() for t in small.predict(q))
    mean0, var0 = oski.ski_posterior(Xs, ys, q, (0.0, 224.0), 30, 1.0, 1.0, 1.0)
    np.testing.assert_allclose(mean, mean0, atol=1e-6)
    np.testing.assert_allclose(var, var0, atol=1e-6)
    gpr.plot_result(pred)
    assert (tmp_path / "weighted_mask" / "predicted_mask_heatmap.png").exists()
