"""network_interpretation_imagenet_b200 — B200 (sm_100a) engine for the perturbation-interpretation hot path
of LiliMeng/network_interpretation_imagenet: mask synthesis -> classifier scoring -> GP surrogate.

Everything numeric runs in libnib.so (hand-written CUDA, C ABI in include/nib.h); these modules are the thin
host mirror of the reference's Python interfaces for that path.  Importing the package does not load the
library; the first compute call does, and fails loudly if it is missing or no B200 is present.
"""
from . import _lib  # noqa: F401
from .masks import KEEP_MUL, REMOVE_MINMAX, MaskSynth, draw_selections, selection_bits, prep_minmax_u8  # noqa: F401
from .classifier import Classifier  # noqa: F401
from .scoring import score  # noqa: F401
from .gp import ActiveMaskGP, GaussianProcessRegressor, expected_improvement, expected_improvement_device  # noqa: F401
from .engine import PerturbationEngine, shard_range, gather_scores  # noqa: F401
from .pipeline import felzenszwalb, img_as_float, segment_image  # noqa: F401

__all__ = ["MaskSynth", "Classifier", "score", "GaussianProcessRegressor", "ActiveMaskGP", "expected_improvement",
           "expected_improvement_device", "PerturbationEngine", "draw_selections", "selection_bits",
           "prep_minmax_u8", "felzenszwalb", "img_as_float", "segment_image", "shard_range", "gather_scores", "KEEP_MUL", "REMOVE_MINMAX"]
