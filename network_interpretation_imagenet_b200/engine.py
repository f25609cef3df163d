"""The whole hot path behind one object: masks -> classifier -> scores (-> GP), sharded over ranks.

Replaces the body of the reference's hot loops (generate_gp_training_data_imagenet.py:221-266,
generate_gp_training_data_mnist.py:203-269, generate_gp_training_data_cifar.py:307-342,
bayesian_active_learning_imagenet.py:178-198): instead of one mask, one H2D copy, one batch-1 forward and
one device sync per iteration, all N masks of an image are synthesised and scored in micro-batches on the
device, and each rank of a torch.distributed job scores a contiguous slice of the masks (masks are
independent work units); one all-gather of (target_prob, top1) per call is the only
collective.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .classifier import Classifier
from .masks import KEEP_MUL, REMOVE_MINMAX, MaskSynth
from .scoring import score


def shard_range(N: int, rank: int, world: int) -> tuple[int, int, int]:
    """Contiguous slice [lo, hi) of N masks for `rank`, and the padded per-rank length (equal on all ranks so a
    single all_gather_into_tensor works).  Mask order is global, so results are identical to a 1-GPU run."""
    per = (N + world - 1) // world
    lo = min(N, rank * per)
    hi = min(N, lo + per)
    return lo, hi, per


def gather_scores(local: torch.Tensor, N: int, group=None) -> torch.Tensor:
    """All-gather a [per, F] per-rank score block into the global [N, F] table (works on gloo/CPU and nccl/CUDA)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local[:N]
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:N]


class PerturbationEngine:
    """refine_ties: relative top-2 margin (top1 - runner-up) / max|logit| below which a bf16-scored mask is
    re-scored by an fp32 copy of the classifier.  bf16 logits agree with the reference's fp32 logits to <= 1e-2
    relative, so a mask whose margin is inside ~3x that band could flip its arg-max; re-scoring exactly those keeps
    top-1 identical to the reference on every mask at a cost proportional to the number of near-ties."""

    def __init__(self, model, image, segments, target: int, mode: int = KEEP_MUL, precision: str = "bf16",
                 max_batch: int = 128, S: int | None = None, device="cuda", group=None, use_graph: bool = False,
                 refine_ties: float | None = None, streams: int = 1):
        _lib.load()
        self.device = torch.device(device)
        self.target = int(target)
        self.mode = mode
        self.group = group
        self.synth = MaskSynth(image, segments, S=S, device=device)
        self.classifier = model if isinstance(model, Classifier) else Classifier.from_torch(
            model, (self.synth.H, self.synth.W), precision=precision, max_batch=max_batch, streams=streams)
        self.refine_ties = refine_ties if self.classifier.precision == "bf16" else None
        self._model_src = None if isinstance(model, Classifier) else model
        self._fp32 = None
        self.refined = 0
        if self.refine_ties is not None and self._model_src is None:
            raise ValueError("refine_ties needs the torch module (to lower an fp32 copy), not a lowered Classifier")
        if use_graph:
            self.classifier.set_graph(True)
        self.rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1

    def _fp32_classifier(self) -> Classifier:
        if self._fp32 is None:
            self._fp32 = Classifier.from_torch(self._model_src, (self.synth.H, self.synth.W), precision="fp32",
                                               max_batch=min(self.classifier.max_batch, 64))
        return self._fp32

    def score_local(self, sel_bits, out: torch.Tensor | None = None):
        """Scores for the given selections on this rank only.  Returns [n, 2] fp32: (target_prob, top1)."""
        d_sel = self.synth.device_bits(sel_bits)
        n = int(d_sel.shape[0])
        if out is None:
            out = torch.zeros(n, 2, dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        logits = self.classifier.forward_masked(self.synth, d_sel, self.mode)
        s = score(logits, self.target)
        out[:n, 0] = s["target_prob"]
        out[:n, 1] = s["top1"].to(torch.float32)   # class ids < 2^24 are exact in fp32
        if self.refine_ties is not None:
            idx = torch.nonzero(s["margin"] < self.refine_ties).flatten()   # one host sync per call
            if idx.numel() > 0:
                self.refined += int(idx.numel())
                lg = self._fp32_classifier().forward_masked(self.synth, d_sel[idx], self.mode)
                s2 = score(lg, self.target)
                out[idx, 0] = s2["target_prob"]
                out[idx, 1] = s2["top1"].to(torch.float32)
        return out

    def score_masks(self, sel_bits) -> dict:
        """All N masks, sharded over the ranks of `group`; every rank returns the full tables."""
        bits = np.ascontiguousarray(sel_bits, dtype=np.uint64)
        N = bits.shape[0]
        lo, hi, per = shard_range(N, self.rank, self.world)
        local = torch.zeros(per, 2, dtype=torch.float32, device=self.device)
        self.score_local(bits[lo:hi], out=local)
        table = gather_scores(local, N, self.group)
        top1 = table[:, 1].to(torch.int32)
        return {"target_prob": table[:, 0], "top1": top1, "correct": (top1 == self.target)}
