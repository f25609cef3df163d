"""CPU: the oracle restatement against fixtures produced by executing the reference's own statements
(tests/golden/make_golden.py) and against the pinned third-party arithmetic (scikit-learn 1.9.0)."""
import os

import numpy as np
import pytest
import torch

from oracle import classifier as ocls
from oracle import gp as ogp
from oracle import masks as om


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=True)


def test_masks_imagenet_bit_exact(golden_dir):
    g = _load(golden_dir, "masks_imagenet.npz")
    x, seg = g["x"], g["segments"]
    u = np.unique(seg)
    rng = om.make_rng(int(g["seed"]))
    for i in range(g["sel"].shape[0]):
        sel = om.draw_window(rng, u)
        assert sel == list(g["sel"][i])
        mask = om.pixel_mask_keep(seg, sel)
        assert np.array_equal(mask, g["mask"][i])
        out = om.apply_keep(x, mask)
        assert out.dtype == np.float32
        assert np.array_equal(out.view(np.uint32), g["out"][i].view(np.uint32))  # incl. the sign of zeros
    assert (np.signbit(g["out"]) & (g["out"] == 0)).any(), "fixture should contain negative zeros"


@pytest.mark.parametrize("name,draw", [("masks_cifar.npz", "cifar"), ("masks_mnist.npz", "mnist")])
def test_masks_remove_minmax_bit_exact(golden_dir, name, draw):
    g = _load(golden_dir, name)
    org, u8 = om.prep_minmax_u8(g["raw"])
    assert np.array_equal(org.view(np.uint32), g["org"].view(np.uint32))
    assert np.array_equal(u8, g["img_u8"])
    seg = g["segments"]
    u = np.unique(seg)
    rng = om.make_rng(int(g["seed"]))
    for i in range(g["sel"].shape[0]):
        sel = om.draw_subset_cifar(rng, u) if draw == "cifar" else om.draw_subset_mnist(rng, u)
        assert sel == list(g["sel"][i])
        mask = om.pixel_mask_remove(seg, sel)
        assert np.array_equal(mask, g["mask"][i])
        out = om.apply_remove_minmax(org, mask)
        assert np.array_equal(out.view(np.uint32), g["out"][i].view(np.uint32))


def test_selection_bits_roundtrip():
    sels = [[0, 5, 63], [64, 65, 99], []]
    Z = om.selection_bits(sels, 100)
    assert Z.shape == (3, 2)
    X = ogp.bits_to_matrix(Z, 100)
    for i, s in enumerate(sels):
        assert sorted(np.nonzero(X[i])[0].tolist()) == sorted(s)


def test_resnet56_restated_matches_reference_module(golden_dir):
    g = _load(golden_dir, "resnet56.npz")
    m = ocls.load_resnet56()
    y = ocls.forward_logits(m, g["x"]).numpy()
    np.testing.assert_allclose(y, g["logits"], rtol=1e-5, atol=1e-5)
    # the known-answer vector recorded in SURVEY.md §4
    kat = np.array([-1.1524, -5.5548, 14.5349, 2.9369, 3.3827, -10.1788, 5.2296, -1.8864, -5.9056, -1.4126])
    np.testing.assert_allclose(y[0], kat, atol=2e-3)


def test_mnist_net_restated_matches_reference_class(golden_dir):
    """mnist_net.npz = the reference's own Classification_Net statements (mnist :72-105) exec'd + the shipped checkpoint."""
    g = _load(golden_dir, "mnist_net.npz")
    m = ocls.load_mnist_net()
    with torch.no_grad():
        x0, x1, x2, p = m(torch.from_numpy(g["x"]))
    np.testing.assert_allclose(p.numpy(), g["pred0"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(x2.numpy(), g["x2"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(x0.mean((2, 3)).numpy(), g["x0_mean"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(x1.mean((2, 3)).numpy(), g["x1_mean"], rtol=1e-5, atol=1e-5)


def test_mnist_checkpoint_loads_strict():
    m = ocls.load_mnist_net()
    x = torch.rand(2, 1, 28, 28, generator=torch.Generator().manual_seed(1))
    x0, x1, x2, p = m(x)
    assert tuple(p.shape) == (2, 10) and tuple(x0.shape) == (2, 32, 28, 28) and tuple(x2.shape) == (2, 128, 7, 7)


def test_expected_improvement_matches_reference_function(golden_dir):
    g = _load(golden_dir, "ei.npz")
    for gib, key in ((True, "neg_ei_max"), (False, "neg_ei_min")):
        got = ogp.expected_improvement(g["mu"], g["sigma"], g["losses"], greater_is_better=gib)
        np.testing.assert_array_equal(np.isnan(got), np.isnan(g[key]))
        ok = ~np.isnan(got)
        np.testing.assert_allclose(got[ok], g[key][ok], rtol=1e-13, atol=1e-15)


def test_gp_restatement_matches_sklearn_fixture(golden_dir):
    g = _load(golden_dir, "gp_sklearn.npz")
    ell = float(g["length_scale"])
    fit = ogp.gp_fit(g["Xt"], g["yt"], ell, alpha=1e-5, normalize_y=True)
    np.testing.assert_allclose(fit["y_mean"], float(g["y_mean"]), rtol=1e-14)
    np.testing.assert_allclose(fit["y_std"], float(g["y_std"]), rtol=1e-14)
    np.testing.assert_allclose(fit["L"], g["L"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(fit["alpha"], g["alpha_vec"], rtol=1e-6, atol=1e-8)
    mu, var, sd = ogp.gp_predict(fit, g["Xq"])
    np.testing.assert_allclose(mu, g["mu"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(sd, g["std"], rtol=1e-6, atol=1e-9)
    for t, l, dl in zip(g["thetas"], g["lml"], g["lml_grad"]):
        lml, grad = ogp.lml_and_grad(g["Xt"], fit["yn"], float(np.exp(t)))
        np.testing.assert_allclose(lml, l, rtol=1e-9)
        np.testing.assert_allclose(grad, dl, rtol=1e-7, atol=1e-9)


def test_gp_restatement_matches_live_sklearn():
    rng = np.random.RandomState(3)
    X = (rng.rand(40, 50) < 0.4).astype(np.float64)
    y = rng.rand(40)
    sk = ogp.sklearn_gp(optimizer=None, n_restarts_optimizer=0, length_scale=3.0).fit(X, y)
    fit = ogp.gp_fit(X, y, 3.0)
    Xq = (rng.rand(16, 50) < 0.4).astype(np.float64)
    mu_s, sd_s = sk.predict(Xq, return_std=True)
    mu, var, sd = ogp.gp_predict(fit, Xq)
    np.testing.assert_allclose(mu, mu_s, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(sd, sd_s, rtol=1e-7, atol=1e-9)


def test_hamming_equals_sqeuclidean_for_binary_masks():
    rng = np.random.RandomState(0)
    sels = [list(rng.choice(49, 20, replace=False)) for _ in range(12)]
    Z = om.selection_bits(sels, 50)
    X = ogp.bits_to_matrix(Z, 50)
    K = ogp.rbf_gram(X, None, 2.5)
    for i in range(12):
        for j in range(12):
            h = bin(int(Z[i, 0]) ^ int(Z[j, 0])).count("1")
            np.testing.assert_allclose(K[i, j], np.exp(-0.5 * h / 2.5 ** 2), rtol=1e-12)
