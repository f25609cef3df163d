"""Small-shape drivers for compute-sanitizer (memcheck / racecheck / synccheck / initcheck) of the hand-written kernels.

    compute-sanitizer --tool memcheck  python tools/sanitize_targets.py masks score conv fused gp
    compute-sanitizer --tool racecheck python tools/sanitize_targets.py masks score
    compute-sanitizer --tool synccheck python tools/sanitize_targets.py conv fused

Each target runs the kernels once on shapes small enough for the sanitizer's 10-100x slowdown and checks the result
against torch, so a clean sanitizer log also means the run computed the right thing.  Logs go to profiles/ (SURVEY.md §5:
the reference has no race detection of any kind)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import network_interpretation_imagenet_b200 as nib  # noqa: E402
from network_interpretation_imagenet_b200 import _lib  # noqa: E402
from network_interpretation_imagenet_b200.classifier import Classifier, _Builder  # noqa: E402


def t_masks():
    rng = np.random.RandomState(0)
    img = rng.rand(3, 32, 32).astype(np.float32) * 255
    seg = (np.arange(32 * 32).reshape(32, 32) // 64) % 16
    sels = nib.draw_selections("cifar", 16, 300, seed=1)
    bits = nib.selection_bits(sels, 16)
    ms = nib.MaskSynth(img, seg, S=16)
    a = ms.synth(bits, nib.REMOVE_MINMAX)
    b, pm = ms.synth(bits, nib.KEEP_MUL, dtype=torch.bfloat16, layout="nhwc", c_stride=4, pad=3, return_pixel_masks=True)
    torch.cuda.synchronize()
    assert a.shape == (300, 3, 32, 32) and b.shape == (300, 38, 38, 4) and pm.shape == (300, 32, 32)
    assert float(b[:, :3].abs().max()) == 0.0 and float(b[:, :, :3].abs().max()) == 0.0   # halo
    heat = ms.heatmap(bits, np.ones(300, np.float32))
    assert heat.shape == (32, 32)
    print("masks ok")


def t_score():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(300, 1000, generator=g).cuda()
    table = torch.zeros(300, 2, device="cuda")
    s = nib.score(logits, 3, table=table)
    lib = _lib.load()
    cap, words = 16, 1
    sel = torch.arange(300, dtype=torch.int64, device="cuda").view(300, 1)
    idx = torch.empty(cap, dtype=torch.int32, device="cuda")
    so = torch.empty(cap, words, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    tot = torch.zeros(2, dtype=torch.int64, device="cuda")
    thr = float(s["margin"].median())
    _lib.check(lib.nib_tie_compact(s["margin"].data_ptr(), 300, thr, sel.data_ptr(), words, cap, idx.data_ptr(), so.data_ptr(),
                                   cnt.data_ptr(), tot.data_ptr(), _lib.stream_handle()), "nib_tie_compact")
    want = torch.nonzero(s["margin"] < thr).flatten()
    assert int(cnt.item()) == want.numel() and torch.equal(idx.long(), want[:cap]) and torch.equal(so.flatten(), want[:cap])
    assert torch.equal(table[:, 0], s["target_prob"]) and torch.equal(table[:, 1], s["top1"].float())
    print("score ok")


def _conv_net(Cin, Cout, k, stride, pad, H, res):
    g = torch.Generator().manual_seed(Cin + Cout + k)
    N = 3
    w = torch.randn(Cout, Cin, k, k, generator=g) / np.sqrt(Cin * k * k)
    bias = torch.randn(Cout, generator=g) * 0.1
    x = torch.randn(N, Cin, H, H, generator=g)
    b = _Builder(_lib.PREC_BF16, N)
    x_in = b.buffer(H, H, Cin, pooled=False)
    Ho = (H + 2 * pad - k) // stride + 1
    out = b.buffer(Ho, Ho, Cout, pooled=False)
    b.conv(x_in, Cin, out, Cout, w, bias, k, stride, pad, relu=True, res=x_in if res else None, res_C=Cout if res else 0)
    feat = b.buffer(1, 1, Cout)
    b.pool(_lib.POOL_AVG, out, Cout, feat, Ho, Ho, 0)
    b.fc(feat, Cout, 4, torch.zeros(4, Cout), None)
    net = Classifier(b, x_in, (Cin, H, H), 4, "bf16", N, taps={"out": out})
    net.forward(x.cuda())
    got = net.read_tap("out", N).cpu()
    bf = lambda t: t.to(torch.bfloat16).float()
    ref = F.conv2d(bf(x), bf(w), bias, stride=stride, padding=pad)
    if res:
        ref = ref + bf(x)
    ref = ref.clamp_min(0)
    assert net.launch_counts()[1] == 1
    assert (got - ref).abs().max().item() <= 1.5e-2 * max(1.0, ref.abs().max().item())


def t_conv():
    _conv_net(64, 64, 3, 1, 1, 12, False)       # conv_tc3<64, BRES>, im2col
    _conv_net(128, 128, 1, 1, 0, 10, True)      # conv_tc3<128, HAS_RES>, tiled
    _conv_net(64, 256, 1, 1, 0, 9, False)       # conv_tc3<256>
    print("conv ok")


def t_fused():
    g = torch.Generator().manual_seed(1)
    N, K1, N1, N2, H = 3, 64, 256, 64, 10
    x = torch.randn(N, N1, H, H, generator=g)
    w0 = torch.randn(K1, N1, 1, 1, generator=g) / N1 ** 0.5
    wc = torch.randn(N1, K1, 1, 1, generator=g) / K1 ** 0.5
    wa = torch.randn(N2, N1, 1, 1, generator=g) / N1 ** 0.5
    b = _Builder(_lib.PREC_BF16, N)
    x_in = b.buffer(H, H, N1, pooled=False)
    h = b.buffer(H, H, K1, pooled=False)
    y = b.buffer(H, H, N1, pooled=False)
    o = b.buffer(H, H, N2, pooled=False)
    b.conv(x_in, N1, h, K1, w0, None, 1, 1, 0, relu=True)
    b.conv(h, K1, y, N1, wc, None, 1, 1, 0, relu=True, res=x_in, res_C=N1)
    b.conv(y, N1, o, N2, wa, None, 1, 1, 0, relu=True)
    feat = b.buffer(1, 1, N2)
    b.pool(_lib.POOL_AVG, o, N2, feat, H, H, 0)
    b.fc(feat, N2, 4, torch.zeros(4, N2), None)
    net = Classifier(b, x_in, (N1, H, H), 4, "bf16", N, taps={"y": y, "o": o})
    net.forward(x.cuda())
    bf = lambda t: t.to(torch.bfloat16).float()
    hr = bf(F.relu(F.conv2d(bf(x), bf(w0))))
    yr = bf(F.relu(F.conv2d(hr, bf(wc)) + bf(x)))
    orf = F.relu(F.conv2d(yr, bf(wa)))
    assert (net.read_tap("o", N).cpu() - orf).abs().max().item() <= 1.5e-2 * max(1.0, orf.abs().max().item())
    print("fused ok")


def t_gp():
    rng = np.random.RandomState(0)
    S, n, m = 50, 200, 70
    sels = [list(rng.choice(S - 1, size=20, replace=False)) for _ in range(n + m)]
    Z = nib.selection_bits(sels, S)
    y = rng.rand(n)
    gp = nib.GaussianProcessRegressor(alpha=1e-5, length_scale=3.0, optimizer=None).fit(Z[:n], y)
    mu, var, sd = gp.predict_device(Z[n:])
    ei, arg = nib.expected_improvement_device(mu, sd, float(y.max()), True)
    lml, g = gp.log_marginal_likelihood(np.log([2.0]), eval_gradient=True)
    torch.cuda.synchronize()
    assert np.isfinite(lml) and np.isfinite(g[0]) and 0 <= int(arg.item()) < m
    print("gp ok")


if __name__ == "__main__":
    for name in sys.argv[1:] or ["masks", "score", "conv", "fused", "gp"]:
        globals()["t_" + name]()
    torch.cuda.synchronize()
    print("all targets done")
