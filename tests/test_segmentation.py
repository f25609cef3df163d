"""CPU: superpixel label maps.  The product's native felzenszwalb (libnib.so, host C++) against the numpy
restatement of scikit-image's algorithm in oracle/segmentation.py, and the restatement's blur stage against
scipy.ndimage.gaussian_filter (the call scikit-image itself makes).  scikit-image is absent: parity is unpinned beyond that."""
import numpy as np
import pytest
from scipy import ndimage as ndi

from oracle import segmentation as oseg


def _smooth_u8(shape, seed, blur):
    rng = np.random.RandomState(seed)
    img = ndi.gaussian_filter(rng.rand(*shape), [blur, blur, 0])
    img = (img - img.min()) / (img.max() - img.min())
    return (img * 255).astype(np.uint8)


def _piecewise_u8(H, W, C, seed):
    """Flat coloured rectangles + mild noise: many exactly-equal (zero) edge costs, the tie case."""
    rng = np.random.RandomState(seed)
    img = np.zeros((H, W, C), np.uint8)
    for _ in range(12):
        y0, x0 = rng.randint(0, H - 4), rng.randint(0, W - 4)
        y1, x1 = rng.randint(y0 + 2, H), rng.randint(x0 + 2, W)
        img[y0:y1, x0:x1] = rng.randint(0, 256, size=C)
    return img


CASES = [("mnist", (28, 28, 1), 5, 2.0), ("cifar", (32, 32, 3), 10, 2.0), ("imagenet", (224, 224, 3), 50, 4.0),
         ("ragged", (37, 61, 3), 7, 1.5)]


@pytest.mark.parametrize("name,shape,min_size,blur", CASES)
def test_blur_stage_is_bit_identical_to_scipy(name, shape, min_size, blur):
    f = oseg.img_as_float(_smooth_u8(shape, 3, blur))
    assert np.array_equal(oseg.gaussian_blur(f, 0.5), ndi.gaussian_filter(f, sigma=[0.5, 0.5, 0]))
    assert np.array_equal(oseg.gaussian_blur(f, 0.8), ndi.gaussian_filter(f, sigma=[0.8, 0.8, 0]))


@pytest.mark.parametrize("name,shape,min_size,blur", CASES)
def test_native_felzenszwalb_equals_oracle(nib, name, shape, min_size, blur):
    """Reference parameters (scale=100, sigma=0.5, per-dataset min_size): identical label maps, pixel for pixel."""
    for seed in (0, 1):
        u8 = _smooth_u8(shape, seed, blur)
        want = oseg.felzenszwalb(oseg.img_as_float(u8), scale=100, sigma=0.5, min_size=min_size)
        got = nib.felzenszwalb(nib.img_as_float(u8), scale=100, sigma=0.5, min_size=min_size)
        assert got.dtype == np.int64 and got.shape == shape[:2]
        assert np.array_equal(got, want)
        assert np.array_equal(nib.segment_image(u8, min_size), want)     # the entry scripts' helper


def test_grey_2d_input_and_flat_regions(nib):
    u8 = _piecewise_u8(48, 40, 3, 5)
    want = oseg.felzenszwalb(oseg.img_as_float(u8), scale=100, sigma=0.5, min_size=10)
    got = nib.felzenszwalb(nib.img_as_float(u8), scale=100, sigma=0.5, min_size=10)
    assert np.array_equal(got, want)
    g = _piecewise_u8(30, 30, 1, 6)[:, :, 0]                               # H x W, no channel axis (mnist :187)
    assert np.array_equal(nib.felzenszwalb(nib.img_as_float(g), 100, 0.5, 5), oseg.felzenszwalb(oseg.img_as_float(g), 100, 0.5, 5))


def test_label_map_properties(nib):
    """What the mask stage relies on: labels contiguous 0..S-1 in raster order of first appearance, every segment
    8-connected and no smaller than min_size, a constant image is one segment."""
    u8 = _smooth_u8((96, 80, 3), 9, 3.0)
    lab = nib.felzenszwalb(nib.img_as_float(u8), scale=100, sigma=0.5, min_size=20)
    S = int(lab.max()) + 1
    assert np.array_equal(np.unique(lab), np.arange(S))
    first = [np.flatnonzero(lab.ravel() == s)[0] for s in range(S)]
    assert first == sorted(first)
    assert np.bincount(lab.ravel()).min() >= 20
    for s in range(S):
        _, n = ndi.label(lab == s, structure=np.ones((3, 3)))
        assert n == 1
    flat = np.full((16, 16, 3), 0.25)
    assert nib.felzenszwalb(flat, 100, 0.5, 5).max() == 0


def test_rejects_bad_arguments(nib):
    with pytest.raises(RuntimeError):
        nib.felzenszwalb(np.zeros((1, 8, 3)), 100, 0.5, 5)
    with pytest.raises(TypeError):
        nib.img_as_float(np.zeros((4, 4), np.float32))
