"""Scoring restated (test infrastructure only)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def score(logits, target: int):
    """top-1 = `.max(1, keepdim=True)[1]` (generate_gp_training_data_imagenet.py:248);
    target prob = `F.softmax(mask_output)[0][label]` (bayesian_active_learning_imagenet.py:196-198);
    max prob = `F.softmax(pred0, dim=1).max(1)` (generate_gp_training_data_mnist.py:250-256);
    correct = pred == target (imagenet :257)."""
    lg = torch.as_tensor(logits, dtype=torch.float32)
    top1 = lg.max(1, keepdim=True)[1][:, 0]
    prob = F.softmax(lg, dim=1)
    return (top1.numpy().astype(np.int32), prob[:, target].numpy(), prob.max(1)[0].numpy(),
            (top1 == target).numpy().astype(np.uint8))
