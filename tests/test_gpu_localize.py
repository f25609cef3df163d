"""GPU: heat map -> uint8 -> bounding box -> threshold search on the device against the reference-executed fixture
(tests/golden/localize.npz), OpenCV (the library the reference calls, utils.py:96-109) and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import classifier as ocls
from oracle import localize as oloc
from oracle import masks as om
from oracle import synthetic

pytestmark = pytest.mark.gpu


def test_heatmap_u8_and_threshold_masks_equal_reference_outputs(nib, golden_dir):
    from network_interpretation_imagenet_b200 import localize as loc
    g = np.load(os.path.join(golden_dir, "localize.npz"))
    S = int(g["segments"].max()) + 1
    ms = nib.MaskSynth(np.zeros((3, 24, 24), np.float32), g["segments"], S=S)
    bits = nib.selection_bits([list(s) for s in g["sels"]], S)
    heat = ms.heatmap(bits, g["labels"].astype(np.float32))
    covered = g["heat"] >= 0
    assert np.array_equal(heat.cpu().numpy()[covered], g["heat"][covered].astype(np.float32))
    assert np.array_equal(loc.heat_to_u8(heat).cpu().numpy(), g["gray"])              # bit-exact uint8 view
    wseg, cover = loc.segment_weights(ms, bits, g["labels"].astype(np.float32))
    values = sorted(set(float(wseg[s]) for s in range(S) if cover[s] > 0))
    assert values == list(g["values"])
    tb = loc.threshold_selection_bits(wseg, cover, values)
    _, pm = ms.synth(tb, nib.KEEP_MUL, return_pixel_masks=True)
    assert np.array_equal(pm.cpu().numpy(), g["new_masks"])                            # generate_new_mask, every threshold


@pytest.mark.parametrize("seed,H,W", [(0, 224, 224), (1, 224, 224), (2, 28, 28), (3, 32, 32), (4, 100, 180)])
def test_bounding_box_equals_opencv(nib, seed, H, W):
    from network_interpretation_imagenet_b200 import localize as loc
    import cv2
    rng = np.random.RandomState(seed)
    gray = np.zeros((H, W), np.uint8)
    for _ in range(rng.randint(1, 12)):                       # random blobs: rectangles, discs, diagonal threads, nested rings
        cy, cx, r = rng.randint(0, H), rng.randint(0, W), rng.randint(1, max(2, min(H, W) // 4))
        kind = rng.randint(0, 4)
        v = int(rng.randint(100, 256))
        if kind == 0:
            gray[max(0, cy - r):cy + r, max(0, cx - r):cx + r] = v
        elif kind == 1:
            cv2.circle(gray, (cx, cy), r, v, -1)
        elif kind == 2:
            cv2.line(gray, (cx, cy), (min(W - 1, cx + r), min(H - 1, cy + r)), v, 1)
        else:
            cv2.circle(gray, (cx, cy), r, v, 2)
            gray[min(H - 1, cy), min(W - 1, cx)] = 255
    gray[rng.rand(H, W) > 0.995] = 255                        # salt noise: many one-pixel components
    for thr in (90, 180, 254):
        want = oloc.bounding_box(gray, thr)
        (x, y, w, h), ncomp, nfg = loc.bounding_box(torch.from_numpy(gray).cuda(), thr)
        assert nfg == int((gray > thr).sum())
        assert (w * h) == want[2] * want[3], (thr, (x, y, w, h), want)
        assert (x, y, w, h) == want, (thr, (x, y, w, h), want)
    (x, y, w, h), ncomp, nfg = loc.bounding_box(torch.zeros(H, W, dtype=torch.uint8, device="cuda"), 10)
    assert (x, y, w, h, ncomp, nfg) == (0, 0, 0, 0, 0, 0)


def test_threshold_search_replays_the_reference_binary_search(nib):
    """validate_mask on the device: all candidate thresholds scored in one batch, then the reference's probe sequence.
    Checked against the oracle's sequential search whose predictor is the oracle's own fp32 forward (ResNet-56, CIFAR)."""
    from network_interpretation_imagenet_b200 import localize as loc
    raw = synthetic.synthetic_image("cifar")
    seg = synthetic.voronoi_labels(32, 32, 20, seed=11)
    model = ocls.load_resnet56()
    target = int(ocls.forward_logits(model, raw[None]).argmax(1)[0])
    eng = nib.PerturbationEngine(model, raw, seg, target, mode=nib.KEEP_MUL, precision="fp32", max_batch=64, S=20)
    sels = nib.draw_selections("window", 20, 60, seed=5)
    bits = nib.selection_bits(sels, 20)
    labels = eng.score_masks(bits)["correct"].to(torch.float32).cpu().numpy()
    res = loc.threshold_search(eng, bits, labels)
    pm = np.stack([np.isin(seg, s).astype(np.uint8) * 255 for s in sels])
    heat, covered = oloc.summed_label_heat(pm, labels)

    def predict_ok(mask):
        x = om.apply_keep(raw, mask)
        return int(ocls.forward_logits(model, x[None]).argmax(1)[0]) == target

    thr, probes, c, w = oloc.threshold_search(heat, covered, predict_ok)
    assert res["threshold"] == thr and res["probes"] == probes
    assert (res["correct_pred_count"], res["wrong_pred_count"]) == (c, w)


def test_summed_heatmap_iou_drop_in(nib):
    """plot_summed_heatmap's numeric tail: device gray map + box + the reference's IOU arithmetic (with and without the
    reference's [x, y, x, y] quirk) against OpenCV / utils on the host."""
    from network_interpretation_imagenet_b200 import localize as loc
    import utils as ru
    rng = np.random.RandomState(7)
    heat = np.zeros((224, 224), np.float32)
    heat[60:150, 40:170] = rng.randint(5, 40, size=(90, 130))
    heat[10:30, 180:220] = 3
    gt = [50, 55, 120, 100]
    iou, box, gray = loc.summed_heatmap_iou(torch.from_numpy(heat).cuda(), 180, gt, quirk=False)
    g0 = oloc.heat_to_u8(heat.astype(np.float64))
    assert np.array_equal(gray.cpu().numpy(), g0)
    x, y, w, h = oloc.bounding_box(g0, 180)
    assert box == [x, y, x + w, y + h]
    assert iou == ru.generate_IOU([x, y, x + w, y + h], [gt[0], gt[1], gt[0] + gt[2], gt[1] + gt[3]])
    iou_q, box_q, _ = loc.summed_heatmap_iou(torch.from_numpy(heat).cuda(), 180, gt, quirk=True)
    assert box_q == [x, y, 2 * x, 2 * y]


def test_display_images_equal_the_reference_arithmetic(nib):
    """./mask_on_img payload: uint8 truncation of the min-max rescaled masked image, fp32 numpy arithmetic in the
    reference's order (bayesian_active_learning_imagenet.py:199-205)."""
    x = synthetic.synthetic_image("imagenet")[:, :64, :96].copy()
    seg = synthetic.voronoi_labels(64, 96, 13, seed=3)
    ms = nib.MaskSynth(x, seg, S=13)
    sels = nib.draw_selections("window", 13, 9, seed=2) + [list(range(13)), []]
    bits = nib.selection_bits(sels, 13)
    got = ms.display_u8(bits, nib.KEEP_MUL).cpu().numpy()
    for i, sel in enumerate(sels):
        mask = np.isin(seg, sel).astype(np.uint8)
        show = (x.copy() * mask).copy().transpose(1, 2, 0)
        show -= show.min()
        with np.errstate(invalid="ignore", divide="ignore"):
            show /= show.max()
        show *= 255
        want = np.nan_to_num(show, nan=0.0).astype(np.uint8)
        assert np.array_equal(got[i], want), i


def test_async_png_writer_round_trip(nib, tmp_path):
    """The side channel: masks and display images leave the device asynchronously and land as the files the GP scripts
    read back (mask_{i}_{label}.png, label parsed from the name: gp_regression.py:66-71)."""
    import cv2
    from network_interpretation_imagenet_b200.pipeline import AsyncPngWriter
    x = synthetic.synthetic_image("cifar")
    seg = synthetic.voronoi_labels(32, 32, 20, seed=11)
    ms = nib.MaskSynth(x, seg, S=20)
    sels = nib.draw_selections("window", 20, 300, seed=1)
    bits = nib.selection_bits(sels, 20)
    _, pm = ms.synth(bits, nib.KEEP_MUL, return_pixel_masks=True)
    labels = np.arange(300) % 2
    w = AsyncPngWriter(workers=4)
    w.submit(str(tmp_path / "masks"), [f"mask_{i}_{labels[i]}.png" for i in range(300)], pm, scale=255, chunk=64)
    w.submit(str(tmp_path / "mask_on_img"), [f"masked_imgs_{i}_{labels[i]}.png" for i in range(300)], ms.display_u8(bits), chunk=64)
    assert w.close() == 600
    for i in (0, 63, 64, 299):
        img = cv2.imread(str(tmp_path / "masks" / f"mask_{i}_{labels[i]}.png"), 0)
        assert np.array_equal(img, np.isin(seg, sels[i]).astype(np.uint8) * 255)
        assert cv2.imread(str(tmp_path / "mask_on_img" / f"masked_imgs_{i}_{labels[i]}.png")).shape == (32, 32, 3)
