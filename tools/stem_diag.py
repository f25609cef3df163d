"""Dumps what the 4-channel stem's TMA box actually puts at each K position: identity-like weights make output channel co
equal to A[row][co] of K block 0, and the input encodes (c, h%8, w%8) in an exactly representable small integer."""
import sys, os, json
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import network_interpretation_imagenet_b200 as nib
from network_interpretation_imagenet_b200 import _lib
from network_interpretation_imagenet_b200.classifier import _Builder, Classifier, _out_hw
H, N, Cout = 32, 1, 64
hh, ww = torch.meshgrid(torch.arange(H), torch.arange(H), indexing="ij")
x = torch.stack([c + 4 * (ww % 8) + 32 * (hh % 8) for c in range(4)]).float()[None]      # [1,4,H,H], values 0..255
w = torch.zeros(Cout, 4, 7, 7)
probe = {}
for co in range(Cout):
    r, s, c = co // 32, (co % 32) // 4, co % 4
    if s < 7:
        w[co, c, r, s] = 1.0
        probe[co] = (c, r, s)
b = _Builder(_lib.PREC_BF16, N)
x_in = b.buffer(H, H, 4, pad=3, pooled=False)
Ho = _out_hw(H, 7, 2, 3)
out = b.buffer(Ho, Ho, Cout, pooled=False)
b.conv(x_in, 4, out, Cout, w, None, 7, 2, 3, relu=False)
feat = b.buffer(1, 1, Cout)
b.pool(_lib.POOL_AVG, out, Cout, feat, Ho, Ho, 0)
b.fc(feat, Cout, 4, torch.zeros(4, Cout), None)
net = Classifier(b, x_in, (4, H, H), 4, "bf16", N, taps={"out": out})
net.forward(x.cuda())
got = net.read_tap("out", N).cpu()[0]          # [64, Ho, Ho]
print("tc launches", net.launch_counts()[1])
def dec(v):
    v = int(round(v)); return (v % 4, (v // 32) % 8, (v // 4) % 8)   # (c, h%8, w%8)
for (p, q) in ((4, 4), (4, 5), (5, 4)):
    row = []
    for co in range(64):
        v = float(got[co, p, q])
        row.append((co, probe.get(co), dec(v) if v == round(v) else v))
    print(f"output pixel p={p} q={q}: expected for probe (c,r,s): (c, (2p+r-3)%8, (2q+s-3)%8)")
    print(row)
