"""Stage 1 host mirror: superpixel-mask selection draws and device mask synthesis.

Mirrors the reference's inlined loop bodies (there is no function boundary in the reference, SURVEY.md
§8b), so the names follow what each block of the loop does:

  selection draws   generate_gp_training_data_imagenet.py:223-231, generate_gp_training_data_mnist.py:208-215,
                    generate_gp_training_data_cifar.py:308, bayesian_active_learning_imagenet.py:173-178
  pixel mask + blend + normalise   imagenet :233-240, mnist :216-242, cifar :310-321, utils.py:92-94
                    -> ONE kernel launch for all N masks (libnib.so nib_mask_synth).

The draws use python's `random` exactly like the reference (`from random import *`, mnist :20-21) so a
seeded run reproduces the reference's mask sequence; they are O(N*k) integer work on the host.
"""
from __future__ import annotations

import ctypes as C
import random as _random
from typing import Iterable, Sequence

import numpy as np
import torch

from . import _lib

KEEP_MUL = _lib.MASK_KEEP_MUL
REMOVE_MINMAX = _lib.MASK_REMOVE_MINMAX


# ---- selection draws (host) --------------------------------------------------------------------
def draw_window(rng: _random.Random, S: int, labels_sorted: Sequence[int] | None = None) -> list[int]:
    """k = int(0.4*S); firstIndex = randint(1, S-k); np.unique(segments)[firstIndex:firstIndex+k]
    (generate_gp_training_data_imagenet.py:223-230)."""
    u = list(range(S)) if labels_sorted is None else list(labels_sorted)
    k = int(0.4 * S)
    f = rng.randint(1, S - k)
    return u[f:f + k]


def window_at(S: int, first_index: int, labels_sorted: Sequence[int] | None = None) -> list[int]:
    """Window at an acquisition-chosen firstIndex (bayesian_active_learning_imagenet.py:173-178)."""
    u = list(range(S)) if labels_sorted is None else list(labels_sorted)
    k = int(0.4 * S)
    return u[first_index:first_index + k]


def draw_subset(rng: _random.Random, S: int, k: int, dummy_randint: bool = False,
                labels_sorted: Sequence[int] | None = None) -> list[int]:
    """sample(range(u[0], u[-1]), k) — the last label is never drawn (mnist :215, cifar :308,
    imagenet :231).  `dummy_randint` replays the unused randint of mnist :211 (it consumes RNG state)."""
    u = list(range(S)) if labels_sorted is None else list(labels_sorted)
    if dummy_randint:
        rng.randint(1, S - k)
    return rng.sample(range(u[0], u[-1]), k)


def draw_selections(mode: str, S: int, N: int, seed: int, k: int | None = None) -> list[list[int]]:
    """N selections in the reference's draw order.  mode: 'window' (imagenet), 'subset_keep' (imagenet :231),
    'mnist' (k=1 removed), 'cifar' (k=5 removed)."""
    rng = _random.Random(seed)
    out = []
    for _ in range(N):
        if mode == "window":
            out.append(draw_window(rng, S))
        elif mode == "subset_keep":
            out.append(draw_subset(rng, S, int(0.4 * S) if k is None else k))
        elif mode == "mnist":
            out.append(draw_subset(rng, S, 1 if k is None else k, dummy_randint=True))
        elif mode == "cifar":
            out.append(draw_subset(rng, S, 5 if k is None else k))
        else:
            raise ValueError(f"unknown selection mode {mode!r}")
    return out


def selection_bits(selections: Iterable[Iterable[int]], S: int) -> np.ndarray:
    """Selection lists -> [N, ceil(S/64)] uint64 bit-vectors (bit s of word s//64 = z[s])."""
    sels = list(selections)
    words = (S + 63) // 64
    out = np.zeros((len(sels), words), dtype=np.uint64)
    for n, sel in enumerate(sels):
        for s in sel:
            if not 0 <= s < S:
                raise ValueError(f"segment id {s} outside [0,{S})")
            out[n, s >> 6] |= np.uint64(1 << (s & 63))
    return out


# ---- device side ---------------------------------------------------------------------------------
def _as_device_labels(segments, S: int, device) -> torch.Tensor:
    seg = torch.as_tensor(np.asarray(segments))
    if seg.min() < 0 or seg.max() >= S:
        raise ValueError("segment labels must lie in [0, S)")
    dt = torch.uint8 if S <= 256 else torch.int16  # int16 carries uint16 payloads up to 32767 labels
    if S > 32767:
        raise ValueError("S > 32767 superpixels is not supported")
    return seg.to(dt).contiguous().to(device)


class MaskSynth:
    """One image + its superpixel label map resident on the device; synthesises masked batches.

    img: C x H x W fp32.  For REMOVE_MINMAX pass the [0,255]-rescaled `org_img`
    (`prep_minmax_u8`, generate_gp_training_data_mnist.py:169-177).
    """

    def __init__(self, img, segments, S: int | None = None, device="cuda"):
        self.lib = _lib.load()
        self.device = torch.device(device)
        img_t = torch.as_tensor(np.asarray(img) if not torch.is_tensor(img) else img, dtype=torch.float32)
        if img_t.dim() != 3:
            raise ValueError("img must be C x H x W")
        self.C, self.H, self.W = (int(v) for v in img_t.shape)
        seg_np = np.asarray(segments.cpu() if torch.is_tensor(segments) else segments)
        if seg_np.shape != (self.H, self.W):
            raise ValueError(f"segments shape {seg_np.shape} != image {(self.H, self.W)}")
        self.S = int(seg_np.max()) + 1 if S is None else int(S)
        self.img = img_t.contiguous().to(self.device)
        self.labels = _as_device_labels(seg_np, self.S, self.device)
        self.label_bytes = self.labels.element_size()
        self.words = (self.S + 63) // 64
        self._seg_minmax = None

    def seg_minmax(self) -> torch.Tensor:
        if self._seg_minmax is None:
            t = torch.empty(self.S, 2, dtype=torch.float32, device=self.device)
            _lib.check(self.lib.nib_segment_minmax(self.img.data_ptr(), self.labels.data_ptr(), self.label_bytes,
                                                   self.C, self.H, self.W, self.S, t.data_ptr(),
                                                   _lib.stream_handle()), "nib_segment_minmax")
            self._seg_minmax = t
        return self._seg_minmax

    def device_bits(self, sel_bits) -> torch.Tensor:
        if torch.is_tensor(sel_bits):
            t = sel_bits
        else:
            a = np.ascontiguousarray(sel_bits, dtype=np.uint64)
            t = torch.from_numpy(a.view(np.int64))
        if t.dim() != 2 or t.shape[1] != self.words:
            raise ValueError(f"selection bits must be [N, {self.words}]")
        return t.contiguous().to(self.device, non_blocking=True)

    def mask_args(self, d_sel: torch.Tensor, mode: int, out: torch.Tensor | None, out_dtype: int, layout: int,
                  c_stride: int = 0, pad: int = 0, pixel_mask: torch.Tensor | None = None) -> _lib.MaskArgs:
        a = _lib.MaskArgs()
        a.d_img = self.img.data_ptr()
        a.d_labels = self.labels.data_ptr()
        a.label_bytes = self.label_bytes
        a.d_sel = d_sel.data_ptr()
        a.sel_words = self.words
        a.N, a.C, a.H, a.W, a.S = int(d_sel.shape[0]), self.C, self.H, self.W, self.S
        a.mode = mode
        a.d_seg_minmax = self.seg_minmax().data_ptr() if mode == REMOVE_MINMAX else None
        a.d_out = out.data_ptr() if out is not None else None
        a.out_dtype = out_dtype
        a.layout = layout
        a.c_stride = c_stride
        a.pad_h = a.pad_w = pad
        a.d_pixel_mask = pixel_mask.data_ptr() if pixel_mask is not None else None
        return a

    def synth(self, sel_bits, mode: int = KEEP_MUL, dtype=torch.float32, layout: str = "nchw", c_stride: int | None = None,
              pad: int = 0, return_pixel_masks: bool = False, out: torch.Tensor | None = None):
        """Masked classifier inputs for all selections: [N,C,H,W] (nchw) or [N,H+2p,W+2p,c_stride] (nhwc)."""
        d_sel = self.device_bits(sel_bits)
        N = int(d_sel.shape[0])
        if layout == "nchw":
            shape = (N, self.C, self.H, self.W)
            lay, cs = _lib.NCHW, 0
        elif layout == "nhwc":
            cs = self.C if c_stride is None else int(c_stride)
            shape = (N, self.H + 2 * pad, self.W + 2 * pad, cs)
            lay = _lib.NHWC
        else:
            raise ValueError(layout)
        if out is None:
            out = torch.empty(shape, dtype=dtype, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != dtype or not out.is_contiguous():
            raise ValueError("out has the wrong shape/dtype")
        pm = torch.empty((N, self.H, self.W), dtype=torch.uint8, device=self.device) if return_pixel_masks else None
        if N == 0:
            return (out, pm) if return_pixel_masks else out
        odt = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}[dtype]
        a = self.mask_args(d_sel, mode, out, odt, lay, cs, pad, pm)
        _lib.check(self.lib.nib_mask_synth(C.byref(a), _lib.stream_handle()), "nib_mask_synth")
        return (out, pm) if return_pixel_masks else out

    def display_u8(self, sel_bits, mode: int = KEEP_MUL) -> torch.Tensor:
        """[N,H,W,C] uint8 display images of the masked inputs, as the reference writes them to ./mask_on_img
        (bayesian_active_learning_imagenet.py:199-216; mnist :225-236 `pic`)."""
        d_sel = self.device_bits(sel_bits)
        N = int(d_sel.shape[0])
        out = torch.empty(N, self.H, self.W, self.C, dtype=torch.uint8, device=self.device)
        if N == 0:
            return out
        a = self.mask_args(d_sel, mode, None, 0, 0)
        a.d_seg_minmax = self.seg_minmax().data_ptr()
        _lib.check(self.lib.nib_mask_display_u8(C.byref(a), out.data_ptr(), _lib.stream_handle()), "nib_mask_display_u8")
        return out

    def heatmap(self, sel_bits, y) -> torch.Tensor:
        """H = sum_i y_i * mask_i over keep-mode masks (gp_regression.py:82-94)."""
        d_sel = self.device_bits(sel_bits)
        yt = torch.as_tensor(y, dtype=torch.float32).contiguous().to(self.device)
        heat = torch.empty(self.H, self.W, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.nib_heatmap(self.labels.data_ptr(), self.label_bytes, self.H, self.W, self.S,
                                        d_sel.data_ptr(), self.words, yt.data_ptr(), int(d_sel.shape[0]),
                                        heat.data_ptr(), _lib.stream_handle()), "nib_heatmap")
        return heat


def prep_minmax_u8(img, device="cuda"):
    """a1 on the device: in-place `x -= min; x /= max; x *= 255` and the truncated uint8 HWC view
    (generate_gp_training_data_mnist.py:169-177).  Returns (org [C,H,W] fp32 cuda, u8 [H,W,C] cuda)."""
    lib = _lib.load()
    t = torch.as_tensor(np.asarray(img) if not torch.is_tensor(img) else img, dtype=torch.float32).clone().contiguous().to(device)
    Cc, H, W = (int(v) for v in t.shape)
    u8 = torch.empty(H, W, Cc, dtype=torch.uint8, device=t.device)
    _lib.check(lib.nib_prep_minmax_u8(t.data_ptr(), Cc, H, W, u8.data_ptr(), _lib.stream_handle()), "nib_prep_minmax_u8")
    return t, u8
