"""CPU: the drop-in utils.py against outputs of the reference's own utils.py functions (tests/golden/utils.npz)."""
import contextlib
import io
import os

import numpy as np

import utils as ours


def test_normalize_image_and_iou_match_reference_outputs(golden_dir):
    g = np.load(os.path.join(golden_dir, "utils.npz"))
    assert np.array_equal(ours.normalize_image(g["img_u8"]), g["norm"])
    with contextlib.redirect_stdout(io.StringIO()):
        iou = np.array([ours.generate_IOU(list(a), list(b)) for a, b in zip(g["boxA"], g["boxB"])])
    assert np.array_equal(iou, g["iou"])          # same integer arithmetic, same division: bit-identical


def test_bounding_box_largest_component_and_reference_quirk(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    gray = np.zeros((64, 64), np.uint8)
    gray[5:15, 8:20] = 200          # 12 x 10
    gray[30:60, 25:45] = 180        # 20 x 30: the largest
    gray[40:42, 2:4] = 90           # below threshold
    assert ours.generate_boundingbox(0, gray, 100, corners=True) == [25, 30, 45, 60]
    assert ours.generate_boundingbox(0, gray, 100) == [25, 30, 25, 30]      # utils.py:109 returns [x, y, x, y]
    assert os.path.exists("heatmaps/gray_img_0.png")
    assert ours.generate_boundingbox(1, np.zeros((8, 8), np.uint8), 100, save=False) == [0, 0, 0, 0]


def test_generate_new_mask_thresholds_the_heat_map():
    heat = np.array([[0.0, 3.0], [7.0, 7.5]])
    assert np.array_equal(ours.generate_new_mask(heat, 7.0), np.array([[0, 0], [0, 1]], np.uint8))
