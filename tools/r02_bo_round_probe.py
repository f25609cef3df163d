"""Where does one acquisition round of BASELINE configs[3] go?  Times posterior / EI / scoring one mask / rank-one append."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import network_interpretation_imagenet_b200 as nib
from network_interpretation_imagenet_b200 import synthetic, gp as _gp
from network_interpretation_imagenet_b200.masks import KEEP_MUL

n = m = 8192
model = synthetic.build_imagenet_model("resnet101")
x = synthetic.synthetic_image("imagenet"); seg = synthetic.voronoi_labels(224, 224, 50)
sels = nib.draw_selections("subset_keep", 50, n + m, seed=1)
bits = nib.selection_bits(sels, 50)
for ties in ("auto", "auto"):
    eng = nib.PerturbationEngine(model, x, seg, target=0, mode=KEEP_MUL, precision="bf16", max_batch=256, S=50, refine_ties=ties)
    y = eng.score_masks(bits[:n])["target_prob"].double().cpu().numpy()
    g = _gp.ActiveMaskGP(bits[n:], alpha=1e-5, length_scale=3.0, normalize_y=True, capacity=32).fit(bits[:n], y)
    acc = {"posterior": 0.0, "ei": 0.0, "score": 0.0, "append": 0.0}
    per_round = []
    def tick(key, fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        acc[key] += dt; per_round.append((key, round(dt * 1e3, 2))); return r
    R = 10
    for it in range(R):
        mu, var, sd = tick("posterior", lambda: g.posterior())
        ei, arg = tick("ei", lambda: _gp.expected_improvement_device(mu, sd, float(np.max(g.y_host)), True))
        j = int(arg.item())
        s = tick("score", lambda: float(eng.score_masks(bits[n + j:n + j + 1])["target_prob"].cpu().numpy()[0]))
        tick("append", lambda: g.append(j, s))
    print(json.dumps({"refine_ties": str(ties), **{k: round(v / R * 1e3, 3) for k, v in acc.items()}}), flush=True)
    print("  per call (ms):", [t for k, t in per_round if k == "posterior"], [t for k, t in per_round if k == "append"], flush=True)
    del eng, g
