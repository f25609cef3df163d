"""Base class of the drop-in `models/*.py` factories: an nn.Module that owns the parameters under the
reference's state-dict names (so `model.load_state_dict(ckpt['state_dict'])` of
generate_gp_training_data_cifar.py:249-250 keeps working, with or without nn.DataParallel) and whose eval-mode
`forward(x)` runs in libnib.so instead of cuDNN.  The lowered network is rebuilt lazily whenever the weights
change.  There is no CPU path: calling it with a CPU tensor, or in training mode, raises."""
from __future__ import annotations

import torch
import torch.nn as nn


class EngineModule(nn.Module):
    precision = "bf16"       # "fp32" for the 1e-4 parity mode
    max_batch = 256
    input_hw = None

    def _invalidate(self, *_):
        self.__dict__["_nib_net"] = None

    def _register_engine_hooks(self):
        self.__dict__["_nib_net"] = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    def configure_engine(self, precision: str | None = None, max_batch: int | None = None):
        if precision is not None:
            self.precision = precision
        if max_batch is not None:
            self.max_batch = max_batch
        self._invalidate()
        return self

    def engine(self, hw=None):
        from network_interpretation_imagenet_b200.classifier import Classifier

        net = self.__dict__.get("_nib_net")
        if net is None:
            net = Classifier.from_torch(self, hw or self.input_hw, precision=self.precision, max_batch=self.max_batch)
            self.__dict__["_nib_net"] = net
        return net

    def _apply(self, fn, *a, **k):   # .cuda()/.float() move the parameters; the lowered copy is stale afterwards
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def train(self, mode: bool = True):
        if mode:
            self._invalidate()
        return super().train(mode)

    def forward(self, x):
        if self.training:
            raise RuntimeError("this drop-in covers the reference's evaluation hot path only (model.eval()); "
                               "training is out of scope (SURVEY.md §2.1)")
        if not x.is_cuda:
            raise RuntimeError("the B200 engine has no CPU path; move the input to CUDA")
        return self.engine(tuple(x.shape[2:])).forward(x)
