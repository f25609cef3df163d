"""CPU: oracle/localize.py against the outputs of the reference's own statements (tests/golden/localize.npz: the function
definitions get_pixel_sorted_mask_label / generate_new_mask and the heat-map uint8 lines, executed by make_golden.py)."""
import os

import numpy as np

from oracle import localize as oloc


def _pixel_masks(g):
    return np.stack([np.isin(g["segments"], sel).astype(np.uint8) * 255 for sel in g["sels"]])


def test_oracle_localize_matches_reference_outputs(golden_dir):
    g = np.load(os.path.join(golden_dir, "localize.npz"))
    heat, covered = oloc.summed_label_heat(_pixel_masks(g), g["labels"])
    assert np.array_equal(covered, g["heat"] >= 0)
    assert np.array_equal(heat[covered], g["heat"][covered])
    assert sorted(set(heat[covered])) == list(g["values"])
    for t, want in zip(g["values"], g["new_masks"]):
        assert np.array_equal(oloc.generate_new_mask(heat, covered, t), want)
    assert np.array_equal(oloc.heat_to_u8(np.where(covered, heat, 0.0)), g["gray"])


def test_oracle_threshold_search_semantics():
    """A monotone classifier (correct iff the mask keeps at least `need` pixels): the search returns the largest threshold
    whose mask is still correct while the next one is not; an always-correct classifier runs off the end (None)."""
    rng = np.random.RandomState(0)
    heat = rng.randint(0, 12, size=(16, 16)).astype(np.float64)
    covered = rng.rand(16, 16) > 0.1
    values = sorted(set(heat[covered]))
    need = int(oloc.generate_new_mask(heat, covered, values[5]).sum())
    thr, probes, c, w = oloc.threshold_search(heat, covered, lambda m: int(m.sum()) >= need)
    assert thr == values[5] and len(probes) <= 5
    thr2, _, _, _ = oloc.threshold_search(heat, covered, lambda m: True)
    assert thr2 is None
