"""torchrun worker for test_two_rank_sharding_matches_single_gpu: every rank must end with the same global
score table as an unsharded run on its own GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import network_interpretation_imagenet_b200 as nib  # noqa: E402
from oracle import classifier as ocls, synthetic  # noqa: E402  (test harness side)

rank = int(os.environ["RANK"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
raw = synthetic.synthetic_image("cifar")
seg = synthetic.voronoi_labels(32, 32, 20, seed=11)
model = ocls.load_resnet56()
d_org, _ = nib.prep_minmax_u8(raw)
sels = nib.draw_selections("cifar", 20, 301, seed=3)   # odd count: exercises the padded last shard
bits = nib.selection_bits(sels, 20)
eng = nib.PerturbationEngine(model, d_org, seg, 3, mode=nib.REMOVE_MINMAX, precision="fp32", max_batch=64, S=20)
out = eng.score_masks(bits)
ref = eng.score_local(bits)        # unsharded on this GPU
assert out["target_prob"].shape[0] == 301
assert torch.equal(out["target_prob"], ref[:, 0]), "sharded and unsharded target probabilities differ"
assert torch.equal(out["top1"], ref[:, 1].to(torch.int32))
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
