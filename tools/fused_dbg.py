"""NIB_TC_DBG=1: role wait timers of one fused expansion+reduction launch at layer-3 shape (K1=256, N1=1024, N2=256)."""
import os, sys
os.environ["NIB_TC_DBG"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from network_interpretation_imagenet_b200 import _lib
from network_interpretation_imagenet_b200.classifier import _Builder, Classifier
K1, N1, N2, H, N = 256, 1024, 256, 14, int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator().manual_seed(0)
b = _Builder(_lib.PREC_BF16, N)
x_in = b.buffer(H, H, N1, pooled=False); h = b.buffer(H, H, K1, pooled=False); y = b.buffer(H, H, N1, pooled=False); o = b.buffer(H, H, N2, pooled=False)
b.conv(x_in, N1, h, K1, torch.randn(K1, N1, 1, 1, generator=g) / 32, None, 1, 1, 0, relu=True)
b.conv(h, K1, y, N1, torch.randn(N1, K1, 1, 1, generator=g) / 16, torch.zeros(N1), 1, 1, 0, relu=True, res=x_in, res_C=N1)
b.conv(y, N1, o, N2, torch.randn(N2, N1, 1, 1, generator=g) / 32, torch.zeros(N2), 1, 1, 0, relu=True)
feat = b.buffer(1, 1, N2); b.pool(_lib.POOL_AVG, o, N2, feat, H, H, 0); b.fc(feat, N2, 4, torch.zeros(4, N2), None)
net = Classifier(b, x_in, (N1, H, H), 4, "bf16", N)
x = torch.randn(N, N1, H, H, generator=g).cuda()
for _ in range(3):
    net.forward(x)
torch.cuda.synchronize()
