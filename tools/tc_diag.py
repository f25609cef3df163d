"""Runs every tcgen05 GEMM / conv case in its own subprocess (a device trap poisons the CUDA context) and writes
one line per case to gpurun_out/tc_diag.txt, so a single GPU call reports on every geometry at once."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASE = r"""
import sys, json, numpy as np, torch
sys.path.insert(0, %(root)r)
import network_interpretation_imagenet_b200 as nib
kind = %(kind)r
if kind == "gemm":
    M, N, K = %(args)r
    lib = nib._lib.load()
    g = torch.Generator().manual_seed(1)
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    B = (torch.randn(N, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    C = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    nib._lib.check(lib.nib_tc_gemm_bf16(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, nib._lib.stream_handle()), "gemm")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    d = (C - ref).abs()
    bad = ~(d <= 1e-3 * max(1.0, ref.abs().max().item()))
    rows = bad.any(1).nonzero().flatten().tolist()
    cols = bad.any(0).nonzero().flatten().tolist()
    print(json.dumps({"max_err": float(torch.nan_to_num(d, nan=1e9).max()), "n_bad": int(bad.sum()), "nan": int(torch.isnan(C).sum()),
                      "bad_rows_head": rows[:16], "bad_cols_head": cols[:16], "n_bad_rows": len(rows), "n_bad_cols": len(cols),
                      "c00": float(C[0, 0]), "ref00": float(ref[0, 0])}))
else:
    sys.path.insert(0, %(root)r + "/tests")
    from test_gpu_classifier import _one_conv_net
    Cin, Cout, k, stride, pad, H, W, relu, residual = %(args)r
    got, ref, net = _one_conv_net(nib, "bf16", Cin, Cout, k, stride, pad, H, W, relu, residual, N=5, seed=3)
    d = (got - ref).abs()
    tol = 1.2e-2 * max(ref.abs().max().item(), 1.0)
    bad = d > tol
    idx = bad.nonzero()
    print(json.dumps({"max_err": float(d.max()), "scale": float(ref.abs().max()), "n_bad": int(bad.sum()), "numel": int(bad.numel()),
                      "first_bad": idx[:6].tolist(), "tc_launches": net.launch_counts()[1],
                      "bad_by_h": bad.any(3).any(1).any(0).nonzero().flatten().tolist()[:20],
                      "bad_by_w": bad.any(2).any(1).any(0).nonzero().flatten().tolist()[:20],
                      "bad_by_n": bad.any(3).any(2).any(1).nonzero().flatten().tolist()}))
"""

GEMMS = [(128, 32, 64), (128, 64, 64), (128, 128, 64), (128, 256, 64), (128, 128, 128), (256, 128, 256), (300, 64, 192),
         (4096, 256, 1024)]
CONVS = [
    (64, 64, 1, 1, 0, 16, 16, False, False), (64, 64, 3, 1, 1, 16, 16, False, False), (64, 64, 3, 1, 1, 14, 14, False, False),
    (128, 128, 3, 2, 1, 28, 28, False, False), (256, 512, 1, 2, 0, 28, 28, False, False), (64, 256, 1, 1, 0, 14, 14, True, True),
    (128, 32, 3, 1, 1, 7, 7, False, False),
    # CTA-pair kernel coverage: every (BLOCK_N, residual) instantiation, multi-tile K, odd tile counts
    (64, 64, 1, 1, 0, 14, 14, True, True), (128, 128, 1, 1, 0, 28, 28, True, True), (256, 1024, 1, 1, 0, 14, 14, True, True),
    (1024, 256, 1, 1, 0, 14, 14, True, False), (256, 256, 3, 1, 1, 14, 14, True, False), (512, 128, 1, 1, 0, 28, 28, True, False),
    (128, 512, 1, 1, 0, 28, 28, True, True), (512, 2048, 1, 1, 0, 7, 7, True, True), (64, 256, 1, 1, 0, 56, 56, True, True),
]


def main():
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    lines = []
    envs = [("", {})]
    for bn in ("32", "64", "256"):
        envs.append((f" [NIB_TC_BLOCK_N={bn}]", {"NIB_TC_BLOCK_N": bn}))
    cases = [("gemm", a, "", {}) for a in GEMMS] + [("conv", a, "", {}) for a in CONVS]
    cases += [("gemm", (256, 256, 256), t, e) for t, e in envs[1:]]
    for kind, a, tag, env in cases:
        code = CASE % {"root": ROOT, "kind": kind, "args": a}
        try:
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180,
                               env={**os.environ, **env})
            tail = (r.stdout.strip().splitlines() or [""])[-1]
            if r.returncode != 0:
                tail = "FAILED rc=%d :: %s" % (r.returncode, (r.stderr.strip().splitlines() or [""])[-1][:300])
        except subprocess.TimeoutExpired:
            tail = "TIMEOUT"
        line = f"{kind} {a}{tag}: {tail}"
        print(line, flush=True)
        lines.append(line)
        with open(os.path.join(out_dir, "tc_diag.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
