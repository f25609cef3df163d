"""Host mirror of the localisation tail of the path: heat map -> threshold search / bounding box / IOU.

Reference call sites (file:line in the reference):
  get_pixel_sorted_mask_label     generate_gp_training_data_imagenet.py:490-515   dict_pixel[p] = sum of labels of the masks covering p
  plot_summed_heatmap             :517-546, bayesian_active_learning_imagenet.py:312-377   min-max to uint8, bbox, IOU
  generate_new_mask               :549-565   mask[p] = 1 where dict_pixel[p] > threshold (covered pixels only)
  validate_mask                   :334-488   binary search over the sorted distinct heat values for the largest threshold whose mask
                                             still gets the right top-1 while the next one does not (two batch-1 forwards per probe)
  generate_boundingbox / IOU      utils.py:96-142

Here the heat map lives on the device (`nib_heatmap`), its 8-bit view and the bounding box are device kernels
(`nib_heat_normalize_u8`, `nib_threshold_bbox`, csrc/localize.cu), and the threshold search scores the masks of ALL
candidate thresholds in one batch through the engine (a heat map built from superpixel masks is constant on superpixels, so
each thresholded mask is again a selection bit-vector over the same label map) and then replays the reference's binary
search on the precomputed predictions: same probes, same decisions, same return value.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def heat_to_u8(heat: torch.Tensor) -> torch.Tensor:
    """uint8((H - H.min()) / (H - H.min()).max() * 255): the reference's float64 numpy arithmetic, on the device."""
    lib = _lib.load()
    h = heat.to(torch.float32).contiguous()
    gray = torch.empty(h.shape, dtype=torch.uint8, device=h.device)
    _lib.check(lib.nib_heat_normalize_u8(h.data_ptr(), h.numel(), gray.data_ptr(), None, _lib.stream_handle()),
               "nib_heat_normalize_u8")
    return gray


def bounding_box(gray: torch.Tensor, threshold: int):
    """(x, y, w, h) of the largest-bounding-box external contour of `gray > threshold` (utils.generate_boundingbox's
    cv2.threshold + findContours + boundingRect), plus (#components, #pixels above the threshold)."""
    lib = _lib.load()
    g = gray.to(torch.uint8).contiguous()
    H, W = (int(v) for v in g.shape)
    box = torch.zeros(6, dtype=torch.int32, device=g.device)
    _lib.check(lib.nib_threshold_bbox(g.data_ptr(), H, W, int(threshold), box.data_ptr(), _lib.stream_handle()),
               "nib_threshold_bbox")
    x, y, w, h, ncomp, nfg = (int(v) for v in box.tolist())
    return (x, y, w, h), ncomp, nfg


def generate_boundingbox(img_index, gray, threshold, corners: bool = False):
    """Device version of utils.generate_boundingbox: returns [x, y, x, y] like the reference (which drops w and h,
    utils.py:109) unless `corners`, then [x, y, x + w, y + h]."""
    (x, y, w, h), _, _ = bounding_box(gray, threshold)
    return [x, y, x + w, y + h] if corners else [x, y, x, y]


def segment_weights(synth, sel_bits, y):
    """(wseg [S] float64, cover [S] int32) on the host: heat value and mask count of every superpixel."""
    lib = _lib.load()
    d_sel = synth.device_bits(sel_bits)
    yt = torch.as_tensor(np.asarray(y, dtype=np.float32)).contiguous().to(synth.device)
    w = torch.empty(synth.S, dtype=torch.float64, device=synth.device)
    c = torch.empty(synth.S, dtype=torch.int32, device=synth.device)
    _lib.check(lib.nib_segment_weights(d_sel.data_ptr(), synth.words, yt.data_ptr(), int(d_sel.shape[0]), synth.S,
                                       w.data_ptr(), c.data_ptr(), _lib.stream_handle()), "nib_segment_weights")
    return w.cpu().numpy(), c.cpu().numpy()


def threshold_selection_bits(wseg: np.ndarray, cover: np.ndarray, thresholds) -> np.ndarray:
    """generate_new_mask for every threshold at once, as selection bit-vectors: segment s is kept iff some mask covered it
    and its summed label exceeds the threshold."""
    S = len(wseg)
    words = (S + 63) // 64
    out = np.zeros((len(thresholds), words), dtype=np.uint64)
    for k, t in enumerate(thresholds):
        for s in range(S):
            if cover[s] > 0 and wseg[s] > t:
                out[k, s >> 6] |= np.uint64(1 << (s & 63))
    return out


def threshold_search(engine, sel_bits, labels, verbose: bool = False):
    """validate_mask's binary search (generate_gp_training_data_imagenet.py:386-478).  `sel_bits` / `labels`: the masks of
    ./masks and the 0/1 label in their file names.  Returns dict(threshold (None if the search ends without the
    correct -> wrong transition), probes, values, correct_pred_count, wrong_pred_count, top1 per candidate threshold)."""
    wseg, cover = segment_weights(engine.synth, sel_bits, labels)
    values = sorted(set(float(wseg[s]) for s in range(len(wseg)) if cover[s] > 0))      # sorted(set(dict_pixel.values()))
    if not values:
        return {"threshold": None, "probes": [], "values": values, "correct_pred_count": 0, "wrong_pred_count": 0, "top1": []}
    bits = threshold_selection_bits(wseg, cover, values)
    top1 = engine.score_masks(bits)["top1"].cpu().numpy()                               # every candidate mask, one batch
    ok = top1 == engine.target
    first, last = 0, len(values) - 1
    probes, correct, wrong, found = [], 0, 0, None
    while first <= last:
        mid = int((first + last) / 2)
        probes.append(mid)
        if mid + 1 >= len(values):
            break        # the reference indexes sorted_dict_values_set[midpoint + 1] here and dies with an IndexError (:398)
        if ok[mid]:
            correct += 1
            if not ok[mid + 1]:
                found = values[mid]
                break
            first = mid + 1
        else:
            wrong += 1
            last = mid - 1
        if verbose:
            print("masked label threshold", values[mid])
    return {"threshold": found, "probes": probes, "values": values, "correct_pred_count": correct, "wrong_pred_count": wrong,
            "top1": top1}


def summed_heatmap_iou(heat: torch.Tensor, bbox_threshold: int, gt_bbox, quirk: bool = True):
    """plot_summed_heatmap's numeric tail (bayesian_active_learning_imagenet.py:349-377): uint8 heat map -> predicted box
    -> IOU with the ground-truth box [x, y, w, h].  `quirk` reproduces the reference exactly: generate_boundingbox returns
    [x, y, x, y] (utils.py:109), to which :369-370 add x and y again; quirk=False uses the real corners [x, y, x+w, y+h]."""
    from utils import generate_IOU
    gray = heat_to_u8(heat)
    pred_box = generate_boundingbox(0, gray, bbox_threshold, corners=not quirk)
    if quirk:
        pred_box[2] += pred_box[0]
        pred_box[3] += pred_box[1]
    gt = [gt_bbox[0], gt_bbox[1], gt_bbox[2] + gt_bbox[0], gt_bbox[3] + gt_bbox[1]]
    return generate_IOU(pred_box, gt), pred_box, gray
