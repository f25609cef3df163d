"""x3 re-score network alone: N live masks through the split-bf16 lowering (for ncu captures of conv_x3_kernel)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import network_interpretation_imagenet_b200 as nib
from network_interpretation_imagenet_b200 import synthetic
from network_interpretation_imagenet_b200.classifier import Classifier
from network_interpretation_imagenet_b200.masks import MaskSynth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
prec = sys.argv[2] if len(sys.argv) > 2 else "x3"
x = synthetic.synthetic_image("imagenet"); seg = synthetic.voronoi_labels(224, 224, 50)
model = synthetic.build_imagenet_model("resnet101")
synth = MaskSynth(x, seg, S=50, device="cuda")
bits = torch.from_numpy(nib.selection_bits(nib.draw_selections("subset_keep", 50, N, seed=1), 50).view(np.int64)).cuda()
net = Classifier.from_torch(model, (224, 224), precision=prec, max_batch=N)
for _ in range(3):
    lg = net.forward_masked(synth, bits, nib.KEEP_MUL)
torch.cuda.synchronize()
prof = net.profile(N)
tot = sum(p[0] for p in prof)
print(f"{prec} N={N}: {tot:.2f} ms per forward (sum of per-op events)")
from collections import defaultdict
g = defaultdict(lambda: [0, 0.0, 0.0])
for ms, kind, fl, geo in prof:
    k = (kind, geo[0], geo[2], geo[3], geo[4], geo[5])
    g[k][0] += 1; g[k][1] += ms; g[k][2] += fl
for k, v in sorted(g.items(), key=lambda kv: -kv[1][1])[:14]:
    print(k, v[0], f"{v[1]:.3f} ms  {v[2] / max(v[1], 1e-9) / 1e9:.1f} TFLOP/s (fp32-grade)")
