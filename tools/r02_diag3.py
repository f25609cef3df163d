"""Round-2 diag: speed and accuracy of the x3 (split-bf16) and fp32 lowerings by live batch size."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import network_interpretation_imagenet_b200 as nib
from network_interpretation_imagenet_b200 import synthetic
from network_interpretation_imagenet_b200.classifier import Classifier
from network_interpretation_imagenet_b200.masks import MaskSynth
arch = sys.argv[1] if len(sys.argv) > 1 else "resnet101"
x = synthetic.synthetic_image("imagenet"); seg = synthetic.voronoi_labels(224, 224, 50)
model = synthetic.build_imagenet_model(arch)
synth = MaskSynth(x, seg, S=50, device="cuda")
bits = torch.from_numpy(nib.selection_bits(nib.draw_selections("subset_keep", 50, 256, seed=1), 50).view(np.int64)).cuda()
out = {"arch": arch}
lib = nib._lib.load()
ref = None
for prec in ("fp32", "x3"):
    net = Classifier.from_torch(model, (224, 224), precision=prec, max_batch=256)
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    nib._lib.check(lib.nib_net_set_dynamic_batch(net.h, cnt.data_ptr()), "dyn")
    lg = torch.empty(256, 1000, device="cuda")
    cost = {}
    for k in (0, 1, 8, 32, 64, 128, 256):
        cnt.fill_(k)
        net.forward_masked(synth, bits, nib.KEEP_MUL, out=lg); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            net.forward_masked(synth, bits, nib.KEEP_MUL, out=lg)
        e1.record(); torch.cuda.synchronize()
        cost[k] = e0.elapsed_time(e1) / 3
    out[prec + "_ms_by_live_masks"] = cost
    if ref is None:
        ref = lg.clone()
    else:
        a, b = lg.double(), ref.double()
        out["x3_vs_fp32_rowwise"] = float(((a - b).abs() / b.abs().amax(1, keepdim=True)).max())
        out["x3_top1_equal"] = bool(torch.equal(a.argmax(1), b.argmax(1)))
    prof = net.profile(64) if False else None
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"r02_diag3_{arch}.json"), "w"), indent=1)
