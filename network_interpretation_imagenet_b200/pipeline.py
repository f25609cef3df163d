"""Shared body of the drop-in entry scripts (generate_gp_training_data_{mnist,cifar,imagenet}.py,
bayesian_active_learning_imagenet.py): everything the reference inlines in its `validate()` / `eval_superpixel()`
hot loops, expressed once over the engine.  Also the on-disk side-channel the GP scripts read
(`./masks/mask_{i}_{label}.png`, gp_regression.py:63-75), written in batches off the device."""
from __future__ import annotations

import os
import shutil

import numpy as np
import torch

from .engine import PerturbationEngine
from .masks import KEEP_MUL, REMOVE_MINMAX, draw_selections, prep_minmax_u8, selection_bits


def felzenszwalb(image: np.ndarray, scale: float = 1.0, sigma: float = 0.8, min_size: int = 20) -> np.ndarray:
    """Drop-in for `skimage.segmentation.felzenszwalb(image, scale, sigma, min_size)` on a float image in [0,1]
    (H x W or H x W x C): libnib's native restatement of scikit-image's algorithm (csrc/segment.cu).  int64 labels H x W,
    contiguous 0..S-1 in raster order of first appearance."""
    import ctypes as C

    from . import _lib
    img = np.ascontiguousarray(np.atleast_3d(np.asarray(image, dtype=np.float64)))
    H, W, Cc = img.shape
    labels = np.empty((H, W), dtype=np.int32)
    S = C.c_int()
    _lib.check(_lib.load().nib_felzenszwalb(img.ctypes.data, H, W, Cc, float(scale), float(sigma), int(min_size),
                                            labels.ctypes.data, C.byref(S)), "nib_felzenszwalb")
    return labels.astype(np.int64)


def img_as_float(img_u8: np.ndarray) -> np.ndarray:
    """`skimage.util.img_as_float` of a uint8 image: float64, multiplied by 1/255 (scikit-image scales unsigned input
    with `np.multiply(image, 1. / imax_in)`, not a division)."""
    if img_u8.dtype != np.uint8:
        raise TypeError("img_as_float restated for uint8 input only (what the reference passes, imagenet :178,:183)")
    return np.multiply(img_u8, 1.0 / 255.0, dtype=np.float64)


def segment_image(img_u8_hwc: np.ndarray, min_size: int, S_fallback: int | None = None, seed: int = 7) -> np.ndarray:
    """`felzenszwalb(img_as_float(img), scale=100, sigma=0.5, min_size=...)` (imagenet :183, cifar :293, mnist :187),
    computed by the library's own implementation (scikit-image is not needed).  S_fallback/seed are accepted for
    backwards compatibility and ignored."""
    return felzenszwalb(img_as_float(np.asarray(img_u8_hwc)), scale=100, sigma=0.5, min_size=min_size)


def reset_dir(path: str) -> None:
    """mask_dir handling of imagenet :207-212: create, or wipe and re-create."""
    if os.path.exists(path):
        shutil.rmtree(path)
    os.makedirs(path)


def write_mask_pngs(mask_dir: str, pixel_masks_u8: np.ndarray, labels, scale: int, start: int = 0) -> None:
    """cv2.imwrite('./masks/mask_{i}_{label}.png', mask*scale)  (imagenet :260/:265 scale=255; mnist/cifar scale=1)."""
    import cv2
    for j, (m, lab) in enumerate(zip(pixel_masks_u8, labels)):
        cv2.imwrite(os.path.join(mask_dir, "mask_{}_{}.png".format(start + j, int(lab))), m * scale)


class AsyncPngWriter:
    """The on-disk side channel (./masks/mask_{i}_{label}.png, ./mask_on_img/...) off the GPU's critical path.

    `submit` copies a batch of uint8 images device -> pinned host on a side stream and returns at once; a pool of
    worker threads waits for the copy's event and runs cv2.imwrite (which releases the GIL), so the stream that scores
    masks never blocks on PNG encoding or the file system.  `close()` drains the queue.  Replaces the synchronous
    cv2.imwrite inside the reference's per-mask loop (generate_gp_training_data_imagenet.py:260-265,
    generate_gp_training_data_mnist.py:263-269, bayesian_active_learning_imagenet.py:207-216)."""

    def __init__(self, workers: int = 8, max_pending: int = 8):
        import queue
        import threading
        self.q = queue.Queue(maxsize=max_pending)
        self.copy_stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self.errors: list[BaseException] = []
        self.written = 0
        self._lock = threading.Lock()
        self.threads = [threading.Thread(target=self._work, daemon=True) for _ in range(workers)]
        for t in self.threads:
            t.start()

    def _work(self):
        import cv2
        while True:
            item = self.q.get()
            if item is None:
                self.q.task_done()
                return
            try:
                ev, host, paths, scale = item
                if ev is not None:
                    ev.synchronize()
                arr = host.numpy()
                for j, p in enumerate(paths):
                    cv2.imwrite(p, arr[j] * scale if scale != 1 else arr[j])
                with self._lock:
                    self.written += len(paths)
            except BaseException as e:   # surfaced by close()
                self.errors.append(e)
            finally:
                self.q.task_done()

    def submit(self, directory: str, names, images_u8: torch.Tensor, scale: int = 1, chunk: int = 256):
        """images_u8: [N,H,W] or [N,H,W,C] uint8 (CUDA or CPU).  One queue item per `chunk` images."""
        names = list(names)
        if images_u8.shape[0] != len(names):
            raise ValueError("one file name per image")
        os.makedirs(directory, exist_ok=True)
        for i in range(0, len(names), chunk):
            part = images_u8[i:i + chunk]
            paths = [os.path.join(directory, n) for n in names[i:i + chunk]]
            if part.is_cuda:
                host = torch.empty(part.shape, dtype=torch.uint8, pin_memory=True)
                self.copy_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self.copy_stream):
                    host.copy_(part, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record()
                part.record_stream(self.copy_stream)
            else:
                host, ev = part.contiguous(), None
            self.q.put((ev, host, paths, scale))

    def close(self):
        self.q.join()
        for _ in self.threads:
            self.q.put(None)
        for t in self.threads:
            t.join()
        if self.errors:
            raise self.errors[0]
        return self.written


def run_generator(kind: str, model, image_chw: np.ndarray, target: int, n_masks: int, mask_seed: int,
                  precision: str = "bf16", max_batch: int = 256, segments: np.ndarray | None = None,
                  mask_dir: str | None = None, S_fallback: int | None = None, verbose: bool = True,
                  mask_on_img_dir: str | None = None):
    """One image through the reference's generator loop.  kind: 'imagenet' (keep window, :221-266),
    'imagenet_subset' (the commented variant :231), 'mnist' (:203-269), 'cifar' (:307-342).
    Returns dict(correct_pred_count, wrong_pred_count, labels u8[N], target_prob f32[N], top1 i32[N], bits, segments)."""
    remove = kind in ("mnist", "cifar")
    if remove:
        d_img, u8 = prep_minmax_u8(image_chw)                     # a1, in place on the device
        u8 = u8.cpu().numpy()
    else:
        disp = image_chw.transpose(1, 2, 0).copy()                # imagenet :171-178 works on a copy
        disp -= disp.min(); disp /= disp.max(); disp *= 255
        u8 = disp.astype(np.uint8)
        d_img = image_chw
    if segments is None:
        defaults = {"imagenet": (50, 50), "imagenet_subset": (50, 50), "cifar": (10, 20), "mnist": (5, 16)}[kind]
        segments = segment_image(u8, defaults[0], S_fallback or defaults[1])
    u = np.unique(segments)
    if not np.array_equal(u, np.arange(len(u))):
        raise ValueError("segment labels must be contiguous 0..S-1")
    S = len(u)
    draw = {"imagenet": "window", "imagenet_subset": "subset_keep", "mnist": "mnist", "cifar": "cifar"}[kind]
    sels = draw_selections(draw, S, n_masks, seed=mask_seed)
    bits = selection_bits(sels, S)
    eng = PerturbationEngine(model, d_img, segments, target, mode=REMOVE_MINMAX if remove else KEEP_MUL,
                             precision=precision, max_batch=max_batch, S=S)
    out = eng.score_masks(bits)
    labels = out["correct"].to(torch.uint8).cpu().numpy()
    res = {"correct_pred_count": int(labels.sum()), "wrong_pred_count": int(n_masks - labels.sum()), "labels": labels,
           "target_prob": out["target_prob"].cpu().numpy(), "top1": out["top1"].cpu().numpy(), "bits": bits,
           "segments": segments, "selections": sels, "engine": eng}
    if mask_dir is not None and eng.rank == 0:
        # ./masks/mask_{i}_{label}.png (+ ./mask_on_img/ when asked): batches leave the device on a side stream and are
        # encoded by worker threads while the next batch is synthesised
        reset_dir(mask_dir)
        writer = AsyncPngWriter()
        top1 = res["top1"]
        for i in range(0, n_masks, 1024):
            sl = slice(i, min(n_masks, i + 1024))
            _, pm = eng.synth.synth(bits[sl], eng.mode, return_pixel_masks=True)
            writer.submit(mask_dir, ["mask_{}_{}.png".format(j, int(labels[j])) for j in range(sl.start, sl.stop)], pm,
                          scale=1 if remove else 255)
            if mask_on_img_dir is not None:
                show = eng.synth.display_u8(bits[sl], eng.mode)
                if show.shape[-1] == 1:
                    show = show[..., 0]
                if remove:   # mnist :263-269 / cifar: masked_imgs_{i}_pred_{pred}_{label}_{prob}.png
                    names = ["masked_imgs_{}_pred_{}_{}_{}.png".format(j, int(top1[j]), int(labels[j]), float(res["target_prob"][j]))
                             for j in range(sl.start, sl.stop)]
                else:        # BO :207-216: masked_imgs_{i}_{label}.png
                    names = ["masked_imgs_{}_{}.png".format(j, int(labels[j])) for j in range(sl.start, sl.stop)]
                writer.submit(mask_on_img_dir, names, show)
        res["png_written"] = writer.close()
    if verbose and eng.rank == 0:
        print("correct_pred_count: ", res["correct_pred_count"])
        print("wrong_pred_count: ", res["wrong_pred_count"])
    return res
