"""Drop-in for the helper functions of the reference's utils.py that sit on / next to the interpretation path.

  normalize_image(image)                       utils.py:92-94   uint8/float [0,255] -> float32 [0,1] by multiplying with 1/255
  generate_boundingbox(img_index, gray, thr)   utils.py:96-109  threshold the heat map, largest external contour's box
  generate_IOU(boxA, boxB, img, idx, folder)   utils.py:114-142 intersection over union of two [x0, y0, x1, y1] boxes
  generate_new_mask(heat, mask_threshold)      generate_gp_training_data_imagenet.py:549-565 pixels whose summed label exceeds
                                               the threshold (on the device heat map instead of the dict_pixel loop)

Reference behaviours kept on purpose: generate_boundingbox returns [x, y, x, y] (the reference drops the
width/height, :109) — pass `corners=True` for [x, y, x + w, y + h]; generate_IOU does not clamp an empty intersection.
The training-side classes of the reference's utils.py (Binarized, Entropy, WeightsCheck, save_checkpoint...) are out of scope.
"""
from __future__ import annotations

import os

import cv2
import numpy as np


def normalize_image(image):
    """Convert pixel intensity values from [0, 255] to [0.0, 1.0]."""
    return np.multiply(image.astype(np.float32), 1.0 / 255.0)


def generate_boundingbox(img_index, gray, threshold, corners: bool = False, save: bool = True):
    """Generate a bounding box for the heatmap"""
    if save:
        os.makedirs("heatmaps", exist_ok=True)
        cv2.imwrite("heatmaps/gray_img_{}.png".format(img_index), gray)
    ret, th1 = cv2.threshold(gray, threshold, 255, cv2.THRESH_BINARY)
    found = cv2.findContours(th1, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    contours = found[-2]            # OpenCV 3 returns (image, contours, hierarchy), OpenCV 4 (contours, hierarchy)
    size = 0
    x, y, w, h = 0, 0, 0, 0
    for contour in contours:
        x_, y_, w_, h_ = cv2.boundingRect(contour)
        if w_ * h_ > size:
            x, y, w, h = x_, y_, w_, h_
            size = w * h
    return [x, y, x + w, y + h] if corners else [x, y, x, y]


def generate_IOU(boxA, boxB, img=None, img_index=0, save_folder=None):
    xA = max(boxA[0], boxB[0])
    yA = max(boxA[1], boxB[1])
    xB = min(boxA[2], boxB[2])
    yB = min(boxA[3], boxB[3])
    interArea = (xB - xA + 1) * (yB - yA + 1)
    print("interArea: ", interArea)
    boxAArea = (boxA[2] - boxA[0] + 1) * (boxA[3] - boxA[1] + 1)
    boxBArea = (boxB[2] - boxB[0] + 1) * (boxB[3] - boxB[1] + 1)
    IOU = interArea / float(boxAArea + boxBArea - interArea)
    if img is not None and save_folder is not None:
        img1 = img.copy()
        img2 = img.copy()
        cv2.rectangle(img1, (int(boxA[0]), int(boxA[1])), (int(boxA[2]), int(boxB[3])), (255, 0, 0), 2)
        cv2.rectangle(img2, (int(boxB[0]), int(boxB[1])), (int(boxB[2]), int(boxB[3])), (0, 0, 255), 2)
        cv2.imwrite(save_folder + "/bbox1_{}.png".format(img_index), img1)
        cv2.imwrite(save_folder + "/bbox2_{}.png".format(img_index), img2)
    return IOU


def generate_new_mask(heat, mask_threshold):
    """result_mask[p] = 1 where the summed label exceeds mask_threshold, else 0 (uint8).  `heat` is the [n, n] heat map
    (CUDA tensor from ski.heatmap_from_masks / MaskSynth.heatmap, or a numpy array); pixels no mask covered hold 0 and
    stay 0, as the reference's `if pixel_pos in dict_pixel` guard leaves them."""
    import torch
    if torch.is_tensor(heat):
        return (heat > mask_threshold).to(torch.uint8)
    return (np.asarray(heat) > mask_threshold).astype(np.uint8)
