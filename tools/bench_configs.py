"""Times BASELINE.json configs[0], [1] and [3] on one B200 through the public package API (configs[2] and [4] are
bench.py's `--arch resnet101` / `--arch densenet121` lines).  One JSON line per config, CUDA-event timed after warm-up.

  configs[0]  generate_gp_training_data_mnist.py: saved MNIST CNN, one synthetic 28x28 image, 256 masks, GP regression fit
  configs[1]  generate_gp_training_data_cifar.py: ResNet-56 checkpoint, 32x32 synthetic image, 4096 masks
  configs[3]  bayesian_active_learning_imagenet.py -a resnet101 --masks: GP posterior on n = 8192 masks, acquisition rounds
"""
import argparse
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
from network_interpretation_imagenet_b200.masks import REMOVE_MINMAX, KEEP_MUL, prep_minmax_u8  # noqa: E402
import BayesianOptimization as bo  # noqa: E402


def ev_time(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def load_ckpt_model(kind):
    if kind == "mnist":
        import generate_gp_training_data_mnist as g
        m = g.Classification_Net()
        p = os.path.join(ROOT, "saved_checkpoints/mnist/checkpoint.pth.tar")
        m.load_state_dict(torch.load(p, map_location="cpu", weights_only=False)["model"])
        return m.eval()
    from importlib import import_module
    m = torch.nn.DataParallel(import_module("models.resnet").createModel(depth=56, data="cifar10", num_classes=10,
                                                                         death_mode="none", death_rate=0.5))
    p = os.path.join(ROOT, "saved_checkpoints/cifar10+-resnet-56/model_best.pth.tar")
    m.load_state_dict(torch.load(p, map_location="cpu", weights_only=False)["state_dict"])
    return m.module.eval()


def generator_config(kind, n_masks, S, precision, max_batch):
    model = load_ckpt_model(kind)
    image = synthetic.synthetic_image(kind)
    H = image.shape[-1]
    seg = synthetic.voronoi_labels(H, H, S, seed=11)
    d_img, _ = prep_minmax_u8(image)
    sels = nib.draw_selections(kind, S, n_masks, seed=3)
    bits = nib.selection_bits(sels, S)
    eng = nib.PerturbationEngine(model, d_img, seg, target=3, mode=REMOVE_MINMAX, precision=precision, max_batch=max_batch, S=S)
    ms = ev_time(lambda: eng.score_masks(bits), 5)
    out = eng.score_masks(bits)
    return eng, bits, out, ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=20)
    ap.add_argument("--n", type=int, default=8192)
    args = ap.parse_args()

    # configs[0]
    eng, bits, out, ms = generator_config("mnist", 256, 16, "fp32", 256)
    y = out["target_prob"].double().cpu().numpy()
    t0 = time.perf_counter()
    gp = nib.GaussianProcessRegressor(alpha=1e-5, n_restarts_optimizer=10, random_state=0)
    gp.fit(bits, y)
    mu, sd = gp.predict(bits, return_std=True)
    torch.cuda.synchronize()
    gp_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"config": "configs[0] mnist: 256 masks + GP regression fit (11 L-BFGS-B starts) + predict",
                      "score_ms": ms, "evals_per_s": 256 / (ms * 1e-3), "gp_fit_predict_ms": gp_ms, "lml_evals": gp.stats["lml_evals"],
                      "length_scale": gp.length_scale_, "precision": "fp32"}), flush=True)
    del eng, gp

    # configs[1]
    for prec in ("fp32", "x3", "bf16"):
        eng, bits, out, ms = generator_config("cifar", 4096, 20, prec, 512)
        print(json.dumps({"config": "configs[1] cifar: ResNet-56 checkpoint, 4096 masks/image", "precision": prec, "score_ms": ms,
                          "evals_per_s": 4096 / (ms * 1e-3), "tcgen05_launches_per_forward": eng.classifier.launch_counts()[1]}), flush=True)
        del eng

    # configs[3]
    n = m = args.n
    model = synthetic.build_imagenet_model("resnet101")
    x = synthetic.synthetic_image("imagenet")
    seg = synthetic.voronoi_labels(224, 224, 50)
    eng = nib.PerturbationEngine(model, x, seg, target=0, mode=KEEP_MUL, precision="bf16", max_batch=256, S=50)
    sels = nib.draw_selections("subset_keep", 50, n + m, seed=1)
    bits = nib.selection_bits(sels, 50)
    # warm-up outside the timed regions (like bench.py's warm-up steps): the first call of the tie policy lowers the
    # re-score network, and the first torch.mv / boolean-index / cuBLAS call of a process costs 100-150 ms each
    # (profiles/r02_bo_round_probe.txt: first posterior() 147 ms, first append() 100 ms, 0.44 / 0.70 ms afterwards)
    eng.score_masks(bits[:4096])
    from network_interpretation_imagenet_b200 import gp as _gpmod
    _w = _gpmod.ActiveMaskGP(bits[n:n + 256], alpha=1e-5, length_scale=3.0, normalize_y=True, capacity=4).fit(bits[:256], np.linspace(0.1, 0.9, 256))
    _w.posterior(); _w.append(0, 0.5); _w.posterior()
    del _w
    torch.cuda.synchronize(); t0 = time.perf_counter()
    y = eng.score_masks(bits[:n])["target_prob"].double().cpu().numpy()
    t_score = time.perf_counter() - t0
    t0 = time.perf_counter()
    Z, yy, hist = bo.bayesian_optimisation_masks(args.rounds, lambda b: eng.score_masks(b)["target_prob"].cpu().numpy(),
                                                 bits[:n], y, bits[n:], length_scale=3.0)
    torch.cuda.synchronize()
    t_bo = time.perf_counter() - t0
    print(json.dumps({"config": f"configs[3] BO: {args.rounds} acquisition rounds, GP posterior on n={n} masks, m={m} candidates (fixed length scale)",
                      "initial_scoring_s": t_score, "initial_evals_per_s": n / t_score, "bo_total_s": t_bo,
                      "ms_per_round": t_bo / args.rounds * 1e3, "picked": [h["candidate"] for h in hist][:5]}), flush=True)


if __name__ == "__main__":
    main()
