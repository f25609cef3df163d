"""Experiment: K classifier replicas (own activation buffers) fed round-robin from K CUDA streams, so an HBM-bound
layer of one micro-batch can overlap a tensor-bound layer of another.  NIB_TC_PAIRS caps each conv kernel's grid."""
import os, sys, json, argparse
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import network_interpretation_imagenet_b200 as nib
from network_interpretation_imagenet_b200 import synthetic
from network_interpretation_imagenet_b200.classifier import Classifier
from network_interpretation_imagenet_b200.masks import MaskSynth
ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=2)
ap.add_argument("--mb", type=int, default=256)
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--steps", type=int, default=4)
a = ap.parse_args()
dev = torch.device("cuda", 0)
x = synthetic.synthetic_image("imagenet"); seg = synthetic.voronoi_labels(224, 224, 50)
model = synthetic.build_imagenet_model("resnet101")
synth = MaskSynth(x, seg, S=50, device=dev)
clfs = [Classifier.from_torch(model, (224, 224), precision="bf16", max_batch=a.mb) for _ in range(a.streams)]
streams = [torch.cuda.Stream() for _ in range(a.streams)]
sels = nib.draw_selections("subset_keep", 50, a.n, seed=1)
bits = torch.from_numpy(nib.selection_bits(sels, 50).view(np.int64)).to(dev)
logits = torch.empty(a.n, 1000, dtype=torch.float32, device=dev)
def step():
    main = torch.cuda.current_stream()
    for s in streams: s.wait_stream(main)
    for j, i in enumerate(range(0, a.n, a.mb)):
        k = j % a.streams
        with torch.cuda.stream(streams[k]):
            clfs[k].forward_masked(synth, bits[i:i + a.mb], nib.KEEP_MUL, out=logits[i:i + a.mb])
    for s in streams: main.wait_stream(s)
for _ in range(2): step()
torch.cuda.synchronize()
ref = logits.clone()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
# single-stream result of replica 0 as the check
chk = clfs[0].forward_masked(synth, bits[:a.mb], nib.KEEP_MUL)
print(json.dumps({"streams": a.streams, "mb": a.mb, "pairs": os.environ.get("NIB_TC_PAIRS"), "serp": os.environ.get("NIB_TC_SERP"),
                  "pdl_off": os.environ.get("NIB_TC_NO_PDL"), "evals_per_s": round(a.n / ms * 1e3), "ms": round(ms, 3),
                  "same": bool(torch.equal(chk, ref[:a.mb]))}))
