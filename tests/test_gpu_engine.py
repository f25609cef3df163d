"""GPU: the whole hot path (masks -> classifier -> scores [-> GP]) against the oracle's restated reference loop."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import classifier as ocls
from oracle import gp as ogp
from oracle import masks as om
from oracle import scoring as oscore
from oracle import synthetic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_loop(model, x, seg, sels, mode, target):
    """One reference loop iteration per mask (batched only for speed; arithmetic per mask is independent)."""
    batch = om.masked_batch(x, seg, sels, mode)
    logits = ocls.forward_logits(model, batch)
    return oscore.score(logits, target), logits.numpy()


def test_mnist_config_end_to_end_fp32(nib):
    """BASELINE configs[0]: saved MNIST CNN, one synthetic 28x28 image, 256 random superpixel masks, GP fit."""
    raw = synthetic.synthetic_image("mnist")
    seg = synthetic.voronoi_labels(28, 28, 16, seed=11)
    org, _ = om.prep_minmax_u8(raw)
    model = ocls.load_mnist_net()
    sels = nib.draw_selections("mnist", 16, 256, seed=42)
    target = int(ocls.forward_logits(model, raw[None]).argmax(1)[0])
    (top1, tprob, mprob, corr), _ = _oracle_loop(model, org, seg, sels, "remove", target)
    d_org, _ = nib.prep_minmax_u8(raw)
    eng = nib.PerturbationEngine(model, d_org, seg, target, mode=nib.REMOVE_MINMAX, precision="fp32", max_batch=128, S=16)
    bits = nib.selection_bits(sels, 16)
    out = eng.score_masks(bits)
    assert np.array_equal(out["top1"].cpu().numpy(), top1)
    np.testing.assert_allclose(out["target_prob"].cpu().numpy(), tprob, rtol=1e-4, atol=1e-6)
    # GP regression over (mask, score) pairs, then posterior on fresh masks: vs the numpy/sklearn oracle
    y = out["target_prob"].cpu().numpy().astype(np.float64)
    uniq = np.unique(bits[:, 0], return_index=True)[1]        # k=1 removal over 15 labels: many duplicates
    gp = nib.GaussianProcessRegressor(alpha=1e-5, length_scale=1.5, optimizer=None).fit(bits[uniq], y[uniq])
    q = nib.selection_bits([[a, b] for a in range(5) for b in range(6, 10)], 16)
    mu, sd = gp.predict(q, return_std=True)
    fit = ogp.gp_fit(ogp.bits_to_matrix(bits[uniq], 16), y[uniq], 1.5)
    mu0, var0, sd0 = ogp.gp_predict(fit, ogp.bits_to_matrix(q, 16))
    assert np.abs(mu - mu0).max() <= 1e-4 and np.abs(sd ** 2 - var0).max() <= 1e-4


def test_cifar_config_end_to_end_fp32(nib):
    raw = synthetic.synthetic_image("cifar")
    seg = synthetic.voronoi_labels(32, 32, 20, seed=11)
    org, _ = om.prep_minmax_u8(raw)
    model = ocls.load_resnet56()
    sels = nib.draw_selections("cifar", 20, 192, seed=7)
    target = int(ocls.forward_logits(model, raw[None]).argmax(1)[0])
    (top1, tprob, _, corr), logits = _oracle_loop(model, org, seg, sels, "remove", target)
    d_org, _ = nib.prep_minmax_u8(raw)
    eng = nib.PerturbationEngine(model, d_org, seg, target, mode=nib.REMOVE_MINMAX, precision="fp32", max_batch=64, S=20)
    out = eng.score_masks(nib.selection_bits(sels, 20))
    assert np.array_equal(out["top1"].cpu().numpy(), top1)
    assert np.array_equal(out["correct"].cpu().numpy().astype(np.uint8), corr)
    np.testing.assert_allclose(out["target_prob"].cpu().numpy(), tprob, rtol=2e-4, atol=1e-6)


def test_stream_copies_give_bit_identical_scores(nib):
    """Classifier copies on side streams (micro-batches round-robin) must not change a single bit of the result."""
    raw = synthetic.synthetic_image("cifar")
    seg = synthetic.voronoi_labels(32, 32, 20, seed=11)
    model = ocls.load_resnet56()
    sels = nib.draw_selections("cifar", 20, 200, seed=3)     # 200 masks, micro-batch 32: ragged last micro-batch
    bits = nib.selection_bits(sels, 20)
    d_org, _ = nib.prep_minmax_u8(raw)
    outs = []
    for streams in (1, 3):
        eng = nib.PerturbationEngine(model, d_org, seg, 0, mode=nib.REMOVE_MINMAX, precision="bf16", max_batch=32, S=20,
                                     streams=streams)
        o = eng.score_masks(bits)
        outs.append((o["target_prob"].cpu().numpy(), o["top1"].cpu().numpy()))
        if streams == 3:
            assert len(eng.classifier._replicas) == 2
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_imagenet_config_end_to_end_bf16(nib):
    """ResNet-101 224^2, S=50 keep-mode masks: identical top-1 on every mask, logits within 1e-2 (bf16)."""
    x = synthetic.synthetic_image("imagenet")
    seg = synthetic.voronoi_labels(224, 224, 50)
    model = ocls.build_imagenet_model("resnet101")
    sels = nib.draw_selections("subset_keep", 50, 24, seed=1)
    target = int(ocls.forward_logits(model, x[None]).argmax(1)[0])
    (top1, tprob, _, _), logits = _oracle_loop(model, x, seg, sels, "keep", target)
    eng = nib.PerturbationEngine(model, x, seg, target, mode=nib.KEEP_MUL, precision="bf16", max_batch=16, S=50)
    assert eng.refine_ties is not None                          # the tie policy is on by default in bf16
    bits = nib.selection_bits(sels, 50)
    got_logits = eng.classifier.forward_masked(eng.synth, bits, nib.KEEP_MUL).cpu().numpy()
    err = (np.abs(got_logits - logits) / np.abs(logits).max(axis=1, keepdims=True)).max()
    assert err <= 1e-2, err
    out = eng.score_masks(bits)
    assert np.array_equal(out["top1"].cpu().numpy(), top1)      # identical top-1 on every mask (ties re-scored in fp32)
    # |d ln p| <= 2 * (logit tolerance) * max|logit|: the bound the 1e-2 logit tolerance implies for a softmax probability
    rtol = float(np.expm1(2 * 1e-2 * np.abs(logits).max()))
    p = out["target_prob"].cpu().numpy().astype(np.float64)
    assert np.all(np.abs(p - tprob) <= rtol * tprob), (float((np.abs(p - tprob) / tprob).max()), rtol)


def test_densenet121_tie_refinement_gives_identical_top1(nib):
    """A seeded random-init DenseNet-121 has near-tied logits (top-2 margin ~1e-2 of max|logit|): pure bf16 flips
    some arg-maxes inside the 1e-2 logit tolerance; with refine_ties those masks are re-scored in fp32."""
    x = synthetic.synthetic_image("imagenet")
    seg = synthetic.voronoi_labels(224, 224, 50)
    model = ocls.build_imagenet_model("densenet121")
    sels = nib.draw_selections("subset_keep", 50, 16, seed=4)
    (top1, tprob, _, _), logits = _oracle_loop(model, x, seg, sels, "keep", 0)
    eng = nib.PerturbationEngine(model, x, seg, 0, mode=nib.KEEP_MUL, precision="bf16", max_batch=16, S=50,
                                 refine_ties=0.05)
    out = eng.score_masks(nib.selection_bits(sels, 50))
    assert np.array_equal(out["top1"].cpu().numpy(), top1), (eng.refined, out["top1"].cpu().numpy(), top1)


def test_cifar_bf16_with_tie_policy_gives_identical_top1_on_every_mask(nib):
    """ResNet-56 in bf16 sits outside the 1e-2 logit tolerance (tests/test_gpu_classifier.py), so its tie band is wider:
    masks whose top-2 margin is inside it are re-scored by the fp32 lowering and top-1 equals the oracle's on all 192."""
    raw = synthetic.synthetic_image("cifar")
    seg = synthetic.voronoi_labels(32, 32, 20, seed=11)
    org, _ = om.prep_minmax_u8(raw)
    model = ocls.load_resnet56()
    sels = nib.draw_selections("cifar", 20, 192, seed=7)
    target = int(ocls.forward_logits(model, raw[None]).argmax(1)[0])
    (top1, tprob, _, corr), logits = _oracle_loop(model, org, seg, sels, "remove", target)
    d_org, _ = nib.prep_minmax_u8(raw)
    eng = nib.PerturbationEngine(model, d_org, seg, target, mode=nib.REMOVE_MINMAX, precision="bf16", max_batch=64, S=20,
                                 refine_ties=0.08, tie_capacity=192)
    out = eng.score_masks(nib.selection_bits(sels, 20))
    st = eng.tie_stats()
    assert st["overflow"] == 0, st
    assert np.array_equal(out["top1"].cpu().numpy(), top1), st
    assert np.array_equal(out["correct"].cpu().numpy().astype(np.uint8), corr)


def test_many_tiny_forwards_on_four_streams_match_one_stream(nib):
    """Stress for scratch shared between streams: 400 masks in micro-batches of 4 over 4 stream copies in
    REMOVE_MINMAX mode - per-mask statistics, segment min/max and the activation buffers of neighbouring micro-batches
    are all in flight at once - must reproduce the single-stream scores bit for bit, several times over."""
    raw = synthetic.synthetic_image("mnist")
    seg = synthetic.voronoi_labels(28, 28, 16, seed=11)
    model = ocls.load_mnist_net()
    sels = nib.draw_selections("mnist", 16, 400, seed=9)
    bits = nib.selection_bits(sels, 16)
    d_org, _ = nib.prep_minmax_u8(raw)
    ref = None
    for streams in (1, 4, 4, 4):
        eng = nib.PerturbationEngine(model, d_org, seg, 3, mode=nib.REMOVE_MINMAX, precision="fp32", max_batch=4, S=16,
                                     streams=streams)
        o = eng.score_masks(bits)
        cur = (o["target_prob"].cpu().numpy(), o["top1"].cpu().numpy())
        if ref is None:
            ref = cur
        assert np.array_equal(ref[0], cur[0]) and np.array_equal(ref[1], cur[1])


def test_two_rank_sharding_matches_single_gpu(nib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(ROOT, "tests", "multi_gpu_worker.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_resnet101_forward_is_bitwise_reproducible(nib):
    """Race hunt for the fused expansion+reduction launches (smem boxes handed from epilogue warps of both CTAs of a pair
    to the tensor core): repeated forwards over two stream copies must reproduce the logits bit for bit."""
    from network_interpretation_imagenet_b200.classifier import Classifier
    from network_interpretation_imagenet_b200.masks import MaskSynth
    x = synthetic.synthetic_image("imagenet")
    seg = synthetic.voronoi_labels(224, 224, 50)
    model = ocls.build_imagenet_model("resnet101")
    synth = MaskSynth(x, seg, S=50, device="cuda")
    sels = nib.draw_selections("subset_keep", 50, 111, seed=5)
    bits = torch.from_numpy(nib.selection_bits(sels, 50).view(np.int64)).cuda()
    clf = Classifier.from_torch(model, (224, 224), precision="bf16", max_batch=37, streams=2)
    ref = clf.forward_masked(synth, bits, nib.KEEP_MUL).clone()
    for _ in range(8):
        assert torch.equal(clf.forward_masked(synth, bits, nib.KEEP_MUL), ref)
