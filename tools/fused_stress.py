"""Race hunt for the fused expansion+reduction kernel: the same ResNet-101 micro-batch forward repeated many times (two stream
copies, different micro-batch sizes) must reproduce its logits bit for bit, and match the unfused path within tolerance."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import network_interpretation_imagenet_b200 as nib
from network_interpretation_imagenet_b200 import synthetic
from network_interpretation_imagenet_b200.classifier import Classifier
from network_interpretation_imagenet_b200.masks import MaskSynth
x = synthetic.synthetic_image("imagenet"); seg = synthetic.voronoi_labels(224, 224, 50)
model = synthetic.build_imagenet_model("resnet101")
synth = MaskSynth(x, seg, S=50, device="cuda")
bad = 0
for mb, n in ((96, 384), (256, 1024), (37, 111)):
    sels = nib.draw_selections("subset_keep", 50, n, seed=mb)
    bits = torch.from_numpy(nib.selection_bits(sels, 50).view(np.int64)).cuda()
    clf = Classifier.from_torch(model, (224, 224), precision="bf16", max_batch=mb, streams=2)
    ref = clf.forward_masked(synth, bits, nib.KEEP_MUL).clone()
    for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 25):
        out = clf.forward_masked(synth, bits, nib.KEEP_MUL)
        if not torch.equal(out, ref):
            bad += 1
            print("MISMATCH", mb, rep, float((out - ref).abs().max()))
    print(f"micro-batch {mb}: {n} masks, logits reproduced; |logit| max {float(ref.abs().max()):.3f}", flush=True)
print("bad =", bad)
sys.exit(1 if bad else 0)
