"""GPU: stage 1 (mask synthesis) through the C ABI — bit-exact against the reference-generated golden
fixtures and against the oracle at the BASELINE configuration sizes."""
import os

import numpy as np
import pytest
import torch

from oracle import masks as om
from oracle import synthetic

pytestmark = pytest.mark.gpu


def _bits_equal(a: torch.Tensor, b: np.ndarray):
    a = a.detach().cpu().numpy()
    assert a.shape == b.shape, (a.shape, b.shape)
    va, vb = a.view(np.uint32), np.ascontiguousarray(b).view(np.uint32)
    nan = np.isnan(a) & np.isnan(b)
    bad = (va != vb) & ~nan
    assert not bad.any(), f"{bad.sum()} of {bad.size} elements differ; first at {np.argwhere(bad)[0]}"


def test_golden_imagenet_keep_mul(nib, golden_dir):
    g = np.load(os.path.join(golden_dir, "masks_imagenet.npz"))
    S = int(g["segments"].max()) + 1
    ms = nib.MaskSynth(g["x"], g["segments"], S=S)
    bits = nib.selection_bits([list(s) for s in g["sel"]], S)
    out, pm = ms.synth(bits, nib.KEEP_MUL, return_pixel_masks=True)
    _bits_equal(out, g["out"])
    assert np.array_equal(pm.cpu().numpy(), g["mask"])


@pytest.mark.parametrize("name", ["masks_cifar.npz", "masks_mnist.npz"])
def test_golden_remove_minmax(nib, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    org, u8 = nib.prep_minmax_u8(g["raw"])
    _bits_equal(org, g["org"])
    assert np.array_equal(u8.cpu().numpy(), g["img_u8"])
    S = int(g["segments"].max()) + 1
    ms = nib.MaskSynth(org, g["segments"], S=S)
    bits = nib.selection_bits([list(s) for s in g["sel"]], S)
    out, pm = ms.synth(bits, nib.REMOVE_MINMAX, return_pixel_masks=True)
    _bits_equal(out, g["out"])
    assert np.array_equal(pm.cpu().numpy(), g["mask"])


@pytest.mark.parametrize("mode", ["window", "subset_keep"])
def test_imagenet_config_vs_oracle(nib, mode):
    x = synthetic.synthetic_image("imagenet")
    seg = synthetic.voronoi_labels(224, 224, 50)
    sels = nib.draw_selections(mode, 50, 48, seed=5)
    want = om.masked_batch(x, seg, sels, "keep")
    ms = nib.MaskSynth(x, seg, S=50)
    bits = nib.selection_bits(sels, 50)
    _bits_equal(ms.synth(bits, nib.KEEP_MUL), want)
    # bf16 NHWC with channel padding and a halo: value = RN-bf16 of the fp32 result, zeros elsewhere
    out = ms.synth(bits, nib.KEEP_MUL, dtype=torch.bfloat16, layout="nhwc", c_stride=8, pad=3)
    assert tuple(out.shape) == (48, 230, 230, 8)
    ref = torch.from_numpy(want).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(out[:, 3:-3, 3:-3, :3].cpu().view(torch.int16), ref.view(torch.int16))
    assert float(out[:, :, :, 3:].abs().max()) == 0.0
    assert float(out[:, :3].abs().max()) == 0.0 and float(out[:, :, -3:].abs().max()) == 0.0
    out4 = ms.synth(bits, nib.KEEP_MUL, dtype=torch.float32, layout="nhwc", c_stride=4)
    assert torch.equal(out4[..., :3].cpu(), torch.from_numpy(want).permute(0, 2, 3, 1))


@pytest.mark.parametrize("kind,draw,S", [("cifar", "cifar", 20), ("mnist", "mnist", 16)])
def test_small_configs_vs_oracle(nib, kind, draw, S):
    raw = synthetic.synthetic_image(kind)
    C, H, W = raw.shape
    seg = synthetic.voronoi_labels(H, W, S, seed=11)
    org, _ = om.prep_minmax_u8(raw)
    sels = nib.draw_selections(draw, S, 256, seed=3)
    want = om.masked_batch(org, seg, sels, "remove")
    d_org, _ = nib.prep_minmax_u8(raw)
    _bits_equal(d_org, org)
    ms = nib.MaskSynth(d_org, seg, S=S)
    _bits_equal(ms.synth(nib.selection_bits(sels, S), nib.REMOVE_MINMAX), want)


def test_many_segments_two_words_and_u16_labels(nib):
    rng = np.random.RandomState(0)
    x = rng.randn(3, 40, 36).astype(np.float32)
    for S in (100, 300):
        seg = synthetic.voronoi_labels(40, 36, S, seed=S)
        sels = [list(rng.choice(S, size=S // 3, replace=False)) for _ in range(9)]
        want = om.masked_batch(x, seg, sels, "keep")
        ms = nib.MaskSynth(x, seg, S=S)
        _bits_equal(ms.synth(nib.selection_bits(sels, S), nib.KEEP_MUL), want)


def test_ragged_and_edge_cases(nib):
    rng = np.random.RandomState(1)
    # W not a multiple of 4 -> scalar path; N = 1; empty and full selections
    x = rng.randn(1, 7, 9).astype(np.float32)
    seg = synthetic.voronoi_labels(7, 9, 5, seed=2)
    sels = [[], [0, 1, 2, 3, 4], [2]]
    ms = nib.MaskSynth(x, seg, S=5)
    _bits_equal(ms.synth(nib.selection_bits(sels, 5), nib.KEEP_MUL), om.masked_batch(x, seg, sels, "keep"))
    _bits_equal(ms.synth(nib.selection_bits(sels[2:], 5), nib.KEEP_MUL), om.masked_batch(x, seg, sels[2:], "keep"))
    empty = ms.synth(np.zeros((0, 1), np.uint64), nib.KEEP_MUL)
    assert tuple(empty.shape) == (0, 1, 7, 9)
    # REMOVE_MINMAX degenerate: everything removed -> 0/0 = NaN in the reference; replicated, not guarded
    org, _ = om.prep_minmax_u8(np.abs(x))
    ms2 = nib.MaskSynth(org, seg, S=5)
    sels2 = [[0, 1, 2, 3, 4], [1], []]
    want = om.masked_batch(org, seg, sels2, "remove")
    assert np.isnan(want[0]).all()
    _bits_equal(ms2.synth(nib.selection_bits(sels2, 5), nib.REMOVE_MINMAX), want)


def test_full_size_properties(nib):
    """N = 2048 masks of 3x224x224 (a rank's share of config 3), checked through size-independent properties."""
    x = synthetic.synthetic_image("imagenet")
    seg = synthetic.voronoi_labels(224, 224, 50)
    ms = nib.MaskSynth(x, seg, S=50)
    sels = nib.draw_selections("subset_keep", 50, 2048, seed=1)
    bits = nib.selection_bits(sels, 50)
    out = ms.synth(bits, nib.KEEP_MUL, dtype=torch.bfloat16, layout="nhwc", c_stride=4)
    xb = torch.from_numpy(x).permute(1, 2, 0).to(torch.bfloat16).cuda()
    # every output pixel is either the image pixel or a zero, decided by the selection bit of its label
    lab = torch.from_numpy(seg).cuda()
    z = torch.from_numpy(bits.view(np.int64)[:, 0].copy()).cuda()
    keep = ((z[:, None, None] >> lab[None]) & 1).bool()
    assert torch.equal(out[..., :3] != 0, keep[..., None] & (xb != 0)[None])
    assert torch.equal(torch.where(keep[..., None], out[..., :3], xb[None].expand_as(out[..., :3])), xb[None].expand_as(out[..., :3]))
    # complement masks partition the image:  x*m + x*(1-m) == x
    comp = (~bits) & np.uint64((1 << 50) - 1)
    a = ms.synth(bits[:64], nib.KEEP_MUL)
    b = ms.synth(comp[:64], nib.KEEP_MUL)
    assert torch.equal(a + b, torch.from_numpy(x).cuda()[None].expand_as(a))
    # kept-pixel count == sum of selected segment areas
    areas = np.bincount(seg.ravel(), minlength=50)
    want = np.array([areas[s].sum() for s in sels[:64]])
    assert np.array_equal(keep[:64].sum(dim=(1, 2)).cpu().numpy(), want)


def test_heatmap_vs_oracle(nib):
    seg = synthetic.voronoi_labels(56, 56, 30, seed=4)
    x = np.zeros((3, 56, 56), np.float32)
    sels = nib.draw_selections("window", 30, 200, seed=8)
    y = (np.arange(200) % 3 == 0).astype(np.float32)
    ms = nib.MaskSynth(x, seg, S=30)
    got = ms.heatmap(nib.selection_bits(sels, 30), y).cpu().numpy()
    np.testing.assert_array_equal(got, om.heatmap(seg, sels, y).astype(np.float32))
