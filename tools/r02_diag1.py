"""Round-2 diagnostics (GPU): (a) margin / bf16-vs-fp32 statistics of the bench workload, (b) fp32 re-score cost,
(c) does a small micro-batch keep the stem / layer-1 / layer-2 activations L2-resident (per-op time per image)."""
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

out = {}
x = synthetic.synthetic_image("imagenet")
seg = synthetic.voronoi_labels(224, 224, 50)
model = synthetic.build_imagenet_model("resnet101")
synth = MaskSynth(x, seg, S=50, device="cuda")
sels = nib.draw_selections("subset_keep", 50, 3072, seed=1)
bits = torch.from_numpy(nib.selection_bits(sels, 50).view(np.int64)).cuda()

# (a) margins + bf16 vs fp32 engine
clf = Classifier.from_torch(model, (224, 224), precision="bf16", max_batch=384, streams=2)
lg = clf.forward_masked(synth, bits, nib.KEEP_MUL)
s = nib.score(lg, 0)
m = s["margin"].cpu().numpy()
out["margin_quantiles"] = {str(q): float(np.quantile(m, q)) for q in (0.001, 0.01, 0.05, 0.1, 0.25, 0.5, 0.9)}
out["n_margin_below"] = {str(t): int((m < t).sum()) for t in (0.002, 0.005, 0.01, 0.02, 0.03, 0.05)}
f32 = Classifier.from_torch(model, (224, 224), precision="fp32", max_batch=64)
torch.cuda.synchronize()
t0 = time.perf_counter()
lg32 = f32.forward_masked(synth, bits[:768], nib.KEEP_MUL)
torch.cuda.synchronize()
out["fp32_768_masks_s"] = time.perf_counter() - t0
for n in (8, 32, 64):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    f32.forward_masked(synth, bits[:n], nib.KEEP_MUL)
    torch.cuda.synchronize()
    out[f"fp32_{n}_masks_ms"] = (time.perf_counter() - t0) * 1e3
a, b = lg[:768].double().cpu().numpy(), lg32.double().cpu().numpy()
mx = np.abs(b).max(axis=1, keepdims=True)
out["bf16_vs_fp32_rel_maxnorm"] = float(np.abs(a - b).max() / np.abs(b).max())
out["bf16_vs_fp32_rel_rowmax"] = float((np.abs(a - b) / mx).max())
out["logit_absmax"] = float(np.abs(b).max())
out["logit_row_std_mean"] = float(b.std(axis=1).mean())
t16, t32 = a.argmax(1), b.argmax(1)
flips = np.nonzero(t16 != t32)[0]
out["top1_flips_of_768"] = int(len(flips))
out["flip_margins"] = [float(m[i]) for i in flips[:20]]
s32 = nib.score(lg32, 0)
out["target_prob_rel_err_max"] = float(((s["target_prob"][:768] - s32["target_prob"]).abs() / s32["target_prob"].abs()).max())
print(json.dumps(out, indent=1), flush=True)
del f32

# (c) per-op time per image at several micro-batches (one stream)
res = {}
for mb in (16, 32, 48, 64, 128, 384):
    c1 = Classifier.from_torch(model, (224, 224), precision="bf16", max_batch=mb, streams=1)
    o = c1.forward_masked(synth, bits[:mb], nib.KEEP_MUL)
    profs = sorted((c1.profile(mb) for _ in range(5)), key=lambda pr: sum(p[0] for p in pr))
    prof = profs[2]
    groups = {"stem+pool": 0.0, "layer1(56)": 0.0, "layer2(28)": 0.0, "layer3(14)": 0.0, "layer4(7)+fc": 0.0}
    for ms, kind, fl, g in prof:
        H = g[0]
        key = "stem+pool" if (H == 112 or (kind == 2 and H == 56)) else "layer1(56)" if H == 56 else "layer2(28)" if H == 28 \
            else "layer3(14)" if H == 14 else "layer4(7)+fc"
        groups[key] += ms
    tot = sum(p[0] for p in prof)
    # whole-forward wall (events) for comparison: 20 back-to-back forwards
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(4, 1536 // mb)
    e0.record()
    for _ in range(reps):
        c1.forward_masked(synth, bits[:mb], nib.KEEP_MUL, out=o)
    e1.record()
    torch.cuda.synchronize()
    res[mb] = {"us_per_image_by_group": {k: v * 1e3 / mb for k, v in groups.items()}, "us_per_image_total_profile": tot * 1e3 / mb,
               "us_per_image_back_to_back": e0.elapsed_time(e1) * 1e3 / (reps * mb)}
    print(mb, json.dumps(res[mb]), flush=True)
    del c1
out["micro_batch_sweep"] = res
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_diag1.json"), "w"), indent=1)
