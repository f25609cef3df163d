"""Round-2 flip study (GPU): over 16384 masks of the bench workload, where do bf16 and fp32 arg-maxes disagree and at
what bf16 margin?  This is the evidence the engine's default tie band rests on (engine.DEFAULT_TIE_BAND)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import network_interpretation_imagenet_b200 as nib  # noqa: E402
from network_interpretation_imagenet_b200 import synthetic  # noqa: E402
from network_interpretation_imagenet_b200.classifier import Classifier  # noqa: E402
from network_interpretation_imagenet_b200.masks import MaskSynth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
arch = sys.argv[2] if len(sys.argv) > 2 else "resnet101"
x = synthetic.synthetic_image("imagenet")
seg = synthetic.voronoi_labels(224, 224, 50)
model = synthetic.build_imagenet_model(arch)
synth = MaskSynth(x, seg, S=50, device="cuda")
sels = nib.draw_selections("subset_keep", 50, N, seed=1)
bits = torch.from_numpy(nib.selection_bits(sels, 50).view(np.int64)).cuda()
clf = Classifier.from_torch(model, (224, 224), precision="bf16", max_batch=384, streams=2)
f32 = Classifier.from_torch(model, (224, 224), precision="fp32", max_batch=128)
out = {"arch": arch, "masks": N}
flips, flip_margin16, flip_margin32 = 0, [], []
max_err, max_diff_err, max_near_err = 0.0, 0.0, 0.0
margins = []
t0 = time.perf_counter()
for i in range(0, N, 2048):
    b = bits[i:i + 2048]
    l16 = clf.forward_masked(synth, b, nib.KEEP_MUL)
    l32 = f32.forward_masked(synth, b, nib.KEEP_MUL)
    s16, s32 = nib.score(l16, 0), nib.score(l32, 0)
    a, r = l16.double(), l32.double()
    amax = r.abs().amax(1, keepdim=True)
    max_err = max(max_err, float(((a - r).abs() / amax).max()))
    t1 = r.argmax(1)
    rows = torch.arange(len(t1), device="cuda")
    de = ((a[rows, t1][:, None] - a) - (r[rows, t1][:, None] - r)).abs() / amax
    max_diff_err = max(max_diff_err, float(de.max()))
    gap32 = (r[rows, t1][:, None] - r) / amax
    near = gap32 < 0.02                                  # only classes this close to the fp32 top-1 can ever overtake it
    max_near_err = max(max_near_err, float(de[near].max()))
    fl = torch.nonzero(s16["top1"] != s32["top1"]).flatten()
    flips += int(fl.numel())
    flip_margin16 += [float(v) for v in s16["margin"][fl]]
    flip_margin32 += [float(v) for v in s32["margin"][fl]]
    margins.append(s16["margin"].cpu().numpy())
torch.cuda.synchronize()
m = np.concatenate(margins)
out.update({"seconds": time.perf_counter() - t0, "flips": flips, "flip_bf16_margins": sorted(flip_margin16),
            "flip_fp32_margins": sorted(flip_margin32), "max_rowwise_logit_err": max_err,
            "max_top1_difference_err_all_classes": max_diff_err, "max_top1_difference_err_near_classes": max_near_err,
            "n_margin_below": {str(t): int((m < t).sum()) for t in (0.001, 0.002, 0.003, 0.004, 0.005, 0.006, 0.008, 0.01, 0.0125, 0.02)}})
# cost model of the fp32 re-score: ms for k live masks in a 128-row buffer (device-side count)
lib = nib._lib.load()
cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
f32b = Classifier.from_torch(model, (224, 224), precision="fp32", max_batch=128)
nib._lib.check(lib.nib_net_set_dynamic_batch(f32b.h, cnt.data_ptr()), "dyn")
lg = torch.empty(128, 1000, device="cuda")
cost = {}
for k in (0, 1, 4, 8, 16, 32, 64, 128):
    cnt.fill_(k)
    f32b.forward_masked(synth, bits[:128], nib.KEEP_MUL, out=lg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        f32b.forward_masked(synth, bits[:128], nib.KEEP_MUL, out=lg)
    e1.record(); torch.cuda.synchronize()
    cost[k] = e0.elapsed_time(e1) / 3
out["fp32_rescore_ms_by_live_masks"] = cost
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"r02_flip_study_{arch}.json"), "w"), indent=1)
