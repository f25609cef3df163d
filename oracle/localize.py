"""Localisation tail of the path restated in numpy / cv2 (test infrastructure only).

Follows generate_gp_training_data_imagenet.py: get_pixel_sorted_mask_label (:490-515), the heat-map min-max / uint8
statements of plot_summed_heatmap (:519-525), generate_new_mask (:549-565), validate_mask's binary search (:386-478), and
utils.py:96-142 (generate_boundingbox / generate_IOU).  Pinned by tests/golden/localize.npz, which holds outputs of the
reference's own statements executed by tests/golden/make_golden.py."""
from __future__ import annotations

import numpy as np


def summed_label_heat(pixel_masks_u8: np.ndarray, labels) -> tuple[np.ndarray, np.ndarray]:
    """(heat [n,n] float64, covered [n,n] bool): dict_pixel[p] = sum of the labels of the masks whose pixel p is 255;
    a pixel is a key of dict_pixel iff some mask covers it (:499-510)."""
    on = pixel_masks_u8 == 255
    heat = (on * np.asarray(labels, dtype=np.float64)[:, None, None]).sum(0)
    return heat, on.any(0)


def heat_to_u8(heat: np.ndarray) -> np.ndarray:
    """:519-525: x = H - H.min(); x = x / x.max(); x *= 255; np.array(x, dtype=np.uint8) in float64."""
    x = np.asarray(heat, dtype=np.float64).copy()
    x = x - x.min()
    with np.errstate(invalid="ignore", divide="ignore"):
        x = x / x.max()
    x *= 255
    return np.array(np.nan_to_num(x, nan=0.0), dtype=np.uint8)


def generate_new_mask(heat: np.ndarray, covered: np.ndarray, mask_threshold) -> np.ndarray:
    """:549-565: 1 where the pixel is a key of dict_pixel and its value exceeds the threshold, else 0."""
    return (covered & (heat > mask_threshold)).astype(np.uint8)


def threshold_search(heat, covered, predict_ok):
    """validate_mask's binary search (:386-478).  predict_ok(mask u8 [n,n]) -> bool (top-1 of the masked image == target).
    Returns (threshold or None, probes, correct_pred_count, wrong_pred_count)."""
    values = sorted(set(float(v) for v in heat[covered]))
    first, last = 0, len(values) - 1
    probes, correct, wrong = [], 0, 0
    while first <= last:
        mid = int((first + last) / 2)
        probes.append(mid)
        if mid + 1 >= len(values):
            return None, probes, correct, wrong          # the reference raises IndexError at values[mid + 1] (:398)
        ok1 = predict_ok(generate_new_mask(heat, covered, values[mid]))
        ok2 = predict_ok(generate_new_mask(heat, covered, values[mid + 1]))
        if ok1:
            correct += 1
            if not ok2:
                return values[mid], probes, correct, wrong
            first = mid + 1
        else:
            wrong += 1
            last = mid - 1
    return None, probes, correct, wrong


def bounding_box(gray_u8: np.ndarray, threshold: int):
    """utils.py:96-109 on OpenCV 4 (2-value findContours): (x, y, w, h) of the contour with the largest w*h."""
    import cv2
    _, th1 = cv2.threshold(gray_u8, threshold, 255, cv2.THRESH_BINARY)
    contours = cv2.findContours(th1, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[-2]
    size = 0
    x, y, w, h = 0, 0, 0, 0
    for contour in contours:
        x_, y_, w_, h_ = cv2.boundingRect(contour)
        if w_ * h_ > size:
            x, y, w, h = x_, y_, w_, h_
            size = w * h
    return x, y, w, h
