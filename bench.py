#!/usr/bin/env python
"""bench.py — masked forward evals/sec of the perturbation-interpretation hot path on N B200s.

Workload (BASELINE.json configs[2], generate_gp_training_data_imagenet.py): ResNet-101 224x224, one synthetic image,
S = 50 superpixels, random keep-masks (the k = int(0.4*S) subset variant of imagenet :231), each mask synthesised,
multiplied into the image, scored through the classifier and reduced to (top-1, target-class probability).
One "step" = `--masks-per-step` masks per GPU through that whole path (weak scaling: per-GPU work is fixed); the
16384-mask job of configs[2] is 8 ranks x 2048 masks = one step at N = 8.  After every step the per-rank score
blocks are all-gathered (NCCL) so each rank holds the global (prob, top1) table, as the GP rank needs.

  value   masks/s with selections already resident in HBM when the timed region starts (CUDA events)
  e2e     same path through the public API with HOST buffers: every step copies the image, label map and
          selection bit-vectors host->device from pinned memory and reads the score table back
  roofline  the dominant kernel (tcgen05 implicit-GEMM conv): algorithmic FLOPs / summed launch time measured
          with CUDA events on the launching stream, against MEASURED_PEAKS.json bf16_tflops_sustained
  cpu_baseline  the oracle's restated reference loop (numpy mask + torch CPU fp32 forward, batch 1, all host threads)

`--impl reference` times that CPU loop alone (the reference has no GPU-independent implementation we could install:
its scripts do not parse on Python 3.12 — see DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "masked forward evals/sec (ResNet-101 224^2)"
UNIT = "evals/s"
FLOPS_PER_EVAL = 15.602810880e9  # SURVEY.md §8d, conv+fc 2*MAC, torchvision 0.26 resnet101
ACT_BYTES_PER_EVAL = (15.78e6 + 16.23e6) * 2  # conv input + output elements per eval, bf16


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--masks-per-step", type=int, default=3072, help="masks per GPU per step")
    ap.add_argument("--micro-batch", type=int, default=384,
                    help="masks per forward; 384 x 196 output pixels = 294 pair tiles = 3.97 waves on 74 CTA pairs in layer 3 "
                         "(256 gives 2.65 waves: same evals/s at the power cap, but 17 %% lower per-kernel rate)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--arch", default="resnet101")
    ap.add_argument("--streams", type=int, default=2,
                    help="copies of the lowered classifier fed round-robin with micro-batches from their own CUDA streams")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gp", action="store_true")
    ap.add_argument("--gp-n", type=int, default=8192, help="GP training-set size (BASELINE configs[3]: n = 8192)")
    ap.add_argument("--graph", action="store_true", help="replay each micro-batch forward as a CUDA graph")
    ap.add_argument("--profile-json", default=None, help="write the per-op CUDA-event profile of one micro-batch here")
    return ap.parse_args()


def ncu_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum of the tcgen05 conv launches of ONE micro-batch-256 forward, from the
    committed ncu capture (profiles/, one pass with cache control on: cold-L2 per launch), averaged per launch."""
    import csv
    p = os.path.join(ROOT, "profiles", "r01_ncu_all_launches_one_forward_v6_mb256.csv")
    if not os.path.exists(p):
        return None, None
    tot, n = 0.0, 0
    for r in csv.DictReader(open(p)):
        if r["kernel"].startswith("conv_tc") or r["kernel"].startswith("conv_fused"):
            tot += float(r["dram__bytes_read.sum"]) + float(r["dram__bytes_write.sum"])
            n += 1
    return (tot / n if n else None), n


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- the CPU arm -----------------------------------------------------------------------------------------
def cpu_reference_loop(model, x, seg, sels, target, torch):
    """One iteration of generate_gp_training_data_imagenet.py:221-266 per mask, restated in oracle/ (batch 1)."""
    from oracle import masks as om
    correct = 0
    with torch.no_grad():
        for sel in sels:
            mask = om.pixel_mask_keep(seg, sel)
            masked = om.apply_keep(x, mask)
            out = model(torch.from_numpy(masked[None]))
            pred = out.max(1, keepdim=True)[1]
            correct += int(pred[0, 0] == target)
    return correct


def cpu_baseline(arch, seconds_budget=12.0):
    import numpy as np
    import torch
    from oracle import classifier as ocls, synthetic
    import network_interpretation_imagenet_b200.masks as pmasks
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = ocls.build_imagenet_model(arch)
    x = synthetic.synthetic_image("imagenet")
    seg = synthetic.voronoi_labels(224, 224, 50)
    sels = pmasks.draw_selections("subset_keep", 50, 4096, seed=1)
    cpu_reference_loop(model, x, seg, sels[:2], 0, torch)            # warm-up
    t0 = time.perf_counter()
    cpu_reference_loop(model, x, seg, sels[2:6], 0, torch)
    per = (time.perf_counter() - t0) / 4
    n = int(max(8, min(1024, seconds_budget / max(per, 1e-4))))
    t0 = time.perf_counter()
    cpu_reference_loop(model, x, seg, sels[6:6 + n], 0, torch)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} masks of the same workload, batch 1 per forward like the reference (imagenet :246), "
                      f"numpy mask ops + torch {torch.__version__} CPU fp32, {cores} threads, {dt:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import classifier as ocls, synthetic
    import network_interpretation_imagenet_b200.masks as pmasks
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = ocls.build_imagenet_model(args.arch)
    x = synthetic.synthetic_image("imagenet")
    seg = synthetic.voronoi_labels(224, 224, 50)
    per_step = 24   # bounded sample of the 2048-mask step: the CPU path is ~10^4 x slower
    total = (args.steps + args.warmup) * per_step
    sels = pmasks.draw_selections("subset_keep", 50, total, seed=1)
    k = 0
    for _ in range(args.warmup):
        cpu_reference_loop(model, x, seg, sels[k:k + per_step], 0, torch); k += per_step
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_loop(model, x, seg, sels[k:k + per_step], 0, torch); k += per_step
    dt = time.perf_counter() - t0
    v = args.steps * per_step / dt
    sample = (f"{per_step} masks per step (bounded sample of the {args.masks_per_step}-mask step), batch 1 per forward, "
              f"oracle port of the reference loop on {cores} host threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "generate_gp_training_data_imagenet.py: ResNet-101 224^2, S=50 superpixels, keep-masks",
                   "masks_per_step_per_gpu": per_step, "arch": args.arch},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ---- our arm ---------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import network_interpretation_imagenet_b200 as nib
    from network_interpretation_imagenet_b200 import _lib
    from network_interpretation_imagenet_b200 import synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    peaks = load_peaks()

    x = synthetic.synthetic_image("imagenet")
    seg = synthetic.voronoi_labels(224, 224, 50)
    model = synthetic.build_imagenet_model(args.arch)     # random-init torchvision resnet101, seeded (no network for weights)
    S, M, mb = 50, args.masks_per_step, args.micro_batch
    eng = nib.PerturbationEngine(model, x, seg, target=0, mode=nib.KEEP_MUL, precision=args.precision, max_batch=mb,
                                 S=S, device=dev, use_graph=args.graph, streams=args.streams)
    clf, synth = eng.classifier, eng.synth
    # global selection table for the whole job, generated identically on every rank from the seed (zero comm)
    total_steps = args.steps + args.warmup
    sels = nib.draw_selections("subset_keep", S, M * world, seed=1)
    bits_host = nib.selection_bits(sels, S)
    lo, hi, per = nib.shard_range(M * world, rank, world)
    my_bits_host = torch.from_numpy(bits_host[lo:hi].view(np.int64)).pin_memory()
    my_bits_dev = my_bits_host.to(dev)
    img_host = torch.from_numpy(x).pin_memory()
    lab_host = torch.from_numpy(seg.astype(np.uint8)).pin_memory()
    scores_host = torch.empty(M * world, 2, dtype=torch.float32).pin_memory()
    local_scores = torch.zeros(per, 2, dtype=torch.float32, device=dev)
    logits = torch.empty(per, clf.num_classes, dtype=torch.float32, device=dev)
    sc = {"top1": torch.empty(per, dtype=torch.int32, device=dev), "target_prob": torch.empty(per, dtype=torch.float32, device=dev),
          "max_prob": torch.empty(per, dtype=torch.float32, device=dev), "correct": torch.empty(per, dtype=torch.uint8, device=dev)}

    def device_step(d_bits):
        """mask synthesis -> classifier (micro-batches of mb, round-robin over the stream copies) -> scores for this
        rank's masks, then the all-gather."""
        n = d_bits.shape[0]
        clf.forward_masked(synth, d_bits, nib.KEEP_MUL, out=logits[:n])
        s = nib.score(logits[:n], 0, out={k: v[:n] for k, v in sc.items()})
        local_scores[:n, 0] = s["target_prob"]
        local_scores[:n, 1] = s["top1"].to(torch.float32)
        return nib.gather_scores(local_scores, M * world)

    def e2e_step():
        synth.img.copy_(img_host, non_blocking=True)
        synth.labels.copy_(lab_host, non_blocking=True)
        d_bits = my_bits_host.to(dev, non_blocking=True)
        table = device_step(d_bits)
        scores_host[: table.shape[0]].copy_(table, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        device_step(my_bits_dev)
    l0, t0 = clf.launch_counts()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda: device_step(my_bits_dev), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    l1, t1 = clf.launch_counts()
    n_mb = (per + mb - 1) // mb
    gpu_launches = (l1 - l0) + args.steps * 1             # + one nib_score kernel per step
    value = M * world * args.steps / (ms / 1e3)

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = M * world * args.steps / (ms_e2e / 1e3)
    h2d = img_host.numel() * 4 + lab_host.numel() + my_bits_host.numel() * 8
    d2h = M * world * 2 * 4

    # roofline of the dominant kernel: per-op CUDA-event timing of one micro-batch forward
    roofline = None
    if rank == 0:
        clf.forward_masked(synth, my_bits_dev[:mb], nib.KEEP_MUL, out=logits[:min(mb, per)])
        profs = sorted((clf.profile(min(mb, per)) for _ in range(5)), key=lambda pr: sum(p[0] for p in pr))
        prof = profs[len(profs) // 2]     # median of five passes by total time (the GPU sits at its power cap: +-5 %)
        if args.profile_json:
            with open(args.profile_json, "w") as f:
                json.dump({"micro_batch": min(mb, per), "ops": [
                    {"ms": p[0], "kind": ["simt_conv", "tc_conv", "pool", "fc"][p[1]], "gflop": p[2] / 1e9,
                     "tflops": (p[2] / (p[0] / 1e3) / 1e12) if p[0] > 0 else None,
                     "Hout": p[3][0], "Wout": p[3][1], "Cin": p[3][2], "Cout": p[3][3], "k": p[3][4], "stride": p[3][5],
                     "residual": p[3][6], "block_n": p[3][7]} for p in prof]}, f, indent=0)
        tc_ms = sum(p[0] for p in prof if p[1] == 1)
        tc_fl = sum(p[2] for p in prof if p[1] == 1)
        all_ms = sum(p[0] for p in prof)
        n_tc = sum(1 for p in prof if p[1] == 1)
        if n_tc and tc_ms > 0:
            ach = tc_fl / (tc_ms / 1e3) / 1e12
            traffic, n_cap = ncu_traffic_per_launch()
            if traffic is not None and min(mb, per) != 256:
                traffic = traffic * min(mb, per) / 256.0    # the capture was taken at micro-batch 256
            roofline = {"bound": "tensor", "kernel": "conv_tc3_kernel + conv_fused_ca_kernel (tcgen05 cta_group::2 implicit GEMM)", "achieved": ach,
                        "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": ach / peaks["bf16_tflops_sustained"], "traffic": traffic,
                        "traffic_source": "profiles/r01_ncu_all_launches_one_forward_v6_mb256.csv: mean DRAM bytes per conv "
                                          "launch (ncu, cold L2 per launch), scaled to this micro-batch",
                        "algorithmic_bytes_per_launch": ACT_BYTES_PER_EVAL * min(mb, per) / n_tc,
                        "flops_per_launch": tc_fl / n_tc,
                        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
                        "launches_per_forward": n_tc, "share_of_forward": tc_ms / all_ms if all_ms else None,
                        # the same flops over (timed-region time x this share): what the conv kernels deliver inside the
                        # real step, where launches run back to back (PDL, two streams) instead of between event records
                        "achieved_in_step": (tc_fl * (per / min(mb, per)) * args.steps / ((ms / 1e3) * (tc_ms / all_ms)) / 1e12)
                        if all_ms else None,
                        "per_kind_ms": {"simt_conv": sum(p[0] for p in prof if p[1] == 0), "tc_conv": tc_ms,
                                        "pool": sum(p[0] for p in prof if p[1] == 2), "fc": sum(p[0] for p in prof if p[1] == 3)}}
        else:
            simt_ms = sum(p[0] for p in prof if p[1] == 0)
            simt_fl = sum(p[2] for p in prof if p[1] == 0)
            ach = simt_fl / (simt_ms / 1e3) / 1e12 if simt_ms > 0 else 0.0
            roofline = {"bound": "tensor", "kernel": "conv_simt_kernel (CUDA cores; tcgen05 path disabled)", "achieved": ach,
                        "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops_sustained"],
                        "traffic": None}

    gp_info = None
    if rank == 0 and not args.no_gp:
        gp_info = gp_bench(nib, torch, np, args.gp_n, S)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.arch)

    if rank == 0:
        line = {
            "metric": METRIC if args.arch == "resnet101" else f"masked forward evals/sec ({args.arch} 224^2)",
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"generate_gp_training_data_imagenet.py: {args.arch} 224^2, S=50 superpixels, "
                                   "k=20 keep-masks, mask synthesis + forward + top-1/softmax scoring + score all-gather",
                       "arch": args.arch, "masks_per_step_per_gpu": M, "micro_batch": mb, "streams": args.streams,
                       "fused_expand_reduce": os.environ.get("NIB_TC_FUSE", "1") != "0", "sharding": f"masks over {world} ranks",
                       "weights": "random init, seeded (no network for pretrained weights)", "cuda_graph": bool(args.graph),
                       "l2": "no explicit flush: each step streams > 10 GB of activations through the 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(gpu_launches), "tcgen05_launches": int(t1 - t0),
            "clocks": clocks, "roofline": roofline,
            "tflops_effective": value / world * FLOPS_PER_EVAL / 1e12 if args.arch == "resnet101" else None,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if gp_info is not None:
            line["gp"] = gp_info
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def gp_bench(nib, torch, np, n, S):
    """GP fit + predict ms at fixed theta on n masks / n candidates (the second half of BASELINE's metric)."""
    rng = np.random.RandomState(0)
    sels = [list(rng.choice(S - 1, size=20, replace=False)) for _ in range(2 * n)]
    Z = nib.selection_bits(sels, S)
    y = rng.rand(n)
    gp = nib.GaussianProcessRegressor(alpha=1e-5, length_scale=3.0, optimizer=None, query_chunk=8192)
    gp.fit(Z[:n], y)
    gp.predict_device(Z[n:])
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    gp.fit(Z[:n], y)
    e[1].record()
    mu, var, sd = gp.predict_device(Z[n:])
    nib.expected_improvement_device(mu, sd, float(y.max()), True)
    e[2].record()
    torch.cuda.synchronize()
    fit_ms, pred_ms = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    chol_flops = n ** 3 / 3
    trsm_flops = float(n) * n * n
    fp64_peak = 37.1   # mma.sync.m8n8k4.f64 chains measured on this pool's B200 (tools/microbench/fp64_peak.cu)
    return {"n": n, "m": n, "fit_ms": fit_ms, "predict_ei_ms": pred_ms, "fit_plus_predict_ms": fit_ms + pred_ms,
            "fp64_tflops_fit": chol_flops / (fit_ms / 1e3) / 1e12, "fp64_tflops_predict": trsm_flops / (pred_ms / 1e3) / 1e12,
            "roofline": {"bound": "fp64 tensor pipe (DMMA)", "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac_fit": chol_flops / (fit_ms / 1e3) / 1e12 / fp64_peak,
                         "frac_predict": trsm_flops / (pred_ms / 1e3) / 1e12 / fp64_peak,
                         "peak_source": "tools/microbench/fp64_peak.cu on B200: DFMA 34.1, DMMA 37.1 TFLOP/s"},
            "note": "fixed theta (optimizer=None); n^3/3 (Cholesky) and m*n^2 (variance TRSM) flops; includes host upload of masks and y"}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
