"""Stage 2 tail host mirror: logits -> (top-1, target-class probability, max probability, correct).

`output.data.max(1, keepdim=True)[1]` (generate_gp_training_data_imagenet.py:248),
`F.softmax(mask_output)[0][label]` (bayesian_active_learning_imagenet.py:196-198),
`F.softmax(pred0, dim=1).max(1)` (generate_gp_training_data_mnist.py:250-256).
"""
from __future__ import annotations

import torch

from . import _lib


def score(logits: torch.Tensor, target: int, out=None, table: torch.Tensor | None = None):
    """logits [N,K] fp32 cuda.  Returns dict(top1 int32[N], target_prob f32[N], max_prob f32[N], correct u8[N],
    margin f32[N] = (top-1 logit - runner-up) / max|logit|, the bf16 tie detector).
    `out` may hold preallocated tensors under the same keys.  `table` ([N,2] fp32 contiguous, optional) receives
    (target_prob, float(top1)) per row from the same kernel: the send buffer of the score all-gather."""
    lib = _lib.load()
    if not logits.is_cuda or logits.dtype != torch.float32 or not logits.is_contiguous():
        raise ValueError("logits must be a contiguous fp32 CUDA tensor")
    N, K = (int(v) for v in logits.shape)
    out = dict(out or {})
    dev = logits.device
    out.setdefault("top1", torch.empty(N, dtype=torch.int32, device=dev))
    out.setdefault("target_prob", torch.empty(N, dtype=torch.float32, device=dev))
    out.setdefault("max_prob", torch.empty(N, dtype=torch.float32, device=dev))
    out.setdefault("correct", torch.empty(N, dtype=torch.uint8, device=dev))
    out.setdefault("margin", torch.empty(N, dtype=torch.float32, device=dev))
    if table is not None and (table.dtype != torch.float32 or tuple(table.shape) != (N, 2) or not table.is_contiguous()):
        raise ValueError("table must be a contiguous [N, 2] fp32 tensor")
    _lib.check(lib.nib_score_table(logits.data_ptr(), N, K, int(target), out["top1"].data_ptr(),
                                   out["target_prob"].data_ptr(), out["max_prob"].data_ptr(), out["correct"].data_ptr(),
                                   out["margin"].data_ptr(), table.data_ptr() if table is not None else None,
                                   _lib.stream_handle()), "nib_score_table")
    return out
