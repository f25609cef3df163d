"""Reference classifiers restated for the CPU oracle (torch fp32 eager).  Test infrastructure only.

State-dict key names match the reference so the shipped checkpoints load strictly:
  - MnistNet      <- Classification_Net, generate_gp_training_data_mnist.py:72-105 (ckpt key 'model', :157-158)
  - ResNetCifar   <- models/resnet.py:10-146 (eval path only: stochastic depth is a no-op when
                     `not self.training`, :31), ckpt saved through nn.DataParallel ('module.' prefix)
  - DenseNetCifar <- models/densenet.py:12-99 with legal module names (the reference's 'norm.1' names
                     raise KeyError on torch >= 1.0; no DenseNet checkpoint is shipped)
  - torchvision resnet101 / densenet121 as loaded at generate_gp_training_data_imagenet.py:579
    (`pretrained=True` needs the network; random init under torch.manual_seed(0) here).
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F


def _conv_bn_relu(i, o, k=3, stride=1, padding=1):
    # generate_gp_training_data_mnist.py:72-77
    return nn.Sequential(nn.Conv2d(i, o, k, stride=stride, padding=padding), nn.BatchNorm2d(o), nn.ReLU(True))


class MnistNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = _conv_bn_relu(1, 32)
        self.conv2 = _conv_bn_relu(32, 32)
        self.conv3 = _conv_bn_relu(32, 64, stride=2)
        self.conv4 = _conv_bn_relu(64, 64)
        self.conv5 = _conv_bn_relu(64, 128, stride=2)
        self.conv6 = nn.Conv2d(128, 128, 3, padding=1)
        self.fc1 = nn.Linear(128, 10)

    def forward(self, x):  # mnist :97-105
        x0 = self.conv2(self.conv1(x))
        x1 = self.conv4(self.conv3(x0))
        x2 = self.conv6(self.conv5(x1))
        f = x2.mean(3).mean(2)
        return x0, x1, x2, self.fc1(f)


class _BasicBlock(nn.Module):
    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample

    def forward(self, x):  # models/resnet.py:26-42, eval branch
        residual = x
        if self.downsample is not None:
            x = self.downsample(x)
        residual = F.relu(self.bn1(self.conv1(residual)))
        residual = self.bn2(self.conv2(residual))
        return F.relu(x + residual)


class _DownsampleB(nn.Module):
    def __init__(self, nIn, nOut, stride):
        super().__init__()
        self.avg = nn.AvgPool2d(stride)
        self.expand_ratio = nOut // nIn

    def forward(self, x):  # models/resnet.py:71-76
        x = self.avg(x)
        return torch.cat([x] + [x.mul(0)] * (self.expand_ratio - 1), 1)


class ResNetCifar(nn.Module):
    def __init__(self, depth=56, num_classes=10):
        super().__init__()
        assert (depth - 2) % 6 == 0
        n = (depth - 2) // 6
        self.inplanes = 16
        self.conv1 = nn.Conv2d(3, 16, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(16)
        self.layer1 = self._make_layer(16, n)
        self.layer2 = self._make_layer(32, n, 2)
        self.layer3 = self._make_layer(64, n, 2)
        self.avgpool = nn.AvgPool2d(8)
        self.fc = nn.Linear(64, num_classes)
        for m in self.modules():  # models/resnet.py:106-112
            if isinstance(m, nn.Conv2d):
                k = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / k))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, planes, blocks, stride=1):
        down = None
        if stride != 1 or self.inplanes != planes:
            down = _DownsampleB(self.inplanes, planes, stride)
        layers = [_BasicBlock(self.inplanes, planes, stride, down)]
        self.inplanes = planes
        layers += [_BasicBlock(planes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        x = self.layer3(self.layer2(self.layer1(x)))
        x = self.avgpool(x)
        return self.fc(x.view(x.size(0), -1))


class _DenseLayer(nn.Module):
    def __init__(self, nin, growth, bn_size):
        super().__init__()
        self.bottleneck = bn_size > 0
        if self.bottleneck:  # models/densenet.py:15-22 (BC)
            self.norm1 = nn.BatchNorm2d(nin)
            self.conv1 = nn.Conv2d(nin, bn_size * growth, 1, bias=False)
            self.norm2 = nn.BatchNorm2d(bn_size * growth)
            self.conv2 = nn.Conv2d(bn_size * growth, growth, 3, padding=1, bias=False)
        else:
            self.norm1 = nn.BatchNorm2d(nin)
            self.conv1 = nn.Conv2d(nin, growth, 3, padding=1, bias=False)

    def forward(self, x):
        y = self.conv1(F.relu(self.norm1(x)))
        if self.bottleneck:
            y = self.conv2(F.relu(self.norm2(y)))
        return torch.cat([x, y], 1)


class DenseNetCifar(nn.Module):
    """models/densenet.py:44-99 for data in {cifar10, cifar100}: 3x3 stem, 3 dense blocks,
    transitions BN-ReLU-1x1-avgpool2 with compression, norm5-ReLU-avg_pool2d(8)-fc."""

    def __init__(self, depth=100, growth_rate=12, num_init_features=24, bn_size=4, compression=0.5, num_classes=10):
        super().__init__()
        n = (depth - 4) // 3
        if bn_size > 0:
            n //= 2
        self.stem = nn.Conv2d(3, num_init_features, 3, padding=1, bias=False)
        nf = num_init_features
        blocks = []
        for b in range(3):
            layers = []
            for _ in range(n):
                layers.append(_DenseLayer(nf, growth_rate, bn_size))
                nf += growth_rate
            blocks.append(nn.Sequential(*layers))
            if b != 2:
                no = int(nf * compression)
                blocks.append(nn.Sequential(nn.BatchNorm2d(nf), nn.ReLU(True), nn.Conv2d(nf, no, 1, bias=False),
                                            nn.AvgPool2d(2, 2)))
                nf = no
        self.blocks = nn.Sequential(*blocks)
        self.norm5 = nn.BatchNorm2d(nf)
        self.classifier = nn.Linear(nf, num_classes)

    def forward(self, x):
        f = F.relu(self.norm5(self.blocks(self.stem(x))))
        f = F.avg_pool2d(f, 8).view(f.size(0), -1)
        return self.classifier(f)


def strip_module_prefix(sd):
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def checkpoints_dir() -> str:
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "saved_checkpoints")


def load_mnist_net() -> MnistNet:
    m = MnistNet()
    ck = torch.load(os.path.join(checkpoints_dir(), "mnist", "checkpoint.pth.tar"), map_location="cpu", weights_only=False)
    m.load_state_dict(ck["model"], strict=True)
    return m.eval()


def load_resnet56() -> ResNetCifar:
    m = ResNetCifar(56, 10)
    ck = torch.load(os.path.join(checkpoints_dir(), "cifar10+-resnet-56", "model_best.pth.tar"), map_location="cpu",
                    weights_only=False)
    m.load_state_dict(strip_module_prefix(ck["state_dict"]), strict=True)
    return m.eval()


def _randomize_bn(model: nn.Module, seed: int):
    """Random-init torchvision nets have BN (gamma,beta,mean,var) = (1,0,0,1): the fold is then a no-op
    and the logits barely depend on the input.  Give BN non-trivial, seeded statistics so the parity
    test exercises the fold and the network output actually varies with the mask."""
    g = torch.Generator().manual_seed(seed)
    for name, m in model.named_modules():
        if isinstance(m, nn.BatchNorm2d):
            m.weight.data = 0.8 + 0.4 * torch.rand(m.num_features, generator=g)
            m.bias.data = 0.1 * torch.randn(m.num_features, generator=g)
            m.running_mean.data = 0.1 * torch.randn(m.num_features, generator=g)
            m.running_var.data = 0.8 + 0.4 * torch.rand(m.num_features, generator=g)
            if name.endswith("bn3"):   # damp the residual branch so a 101-layer random net stays O(1)
                m.weight.data *= 0.25


def build_imagenet_model(arch: str = "resnet101", seed: int = 0, randomize_bn: bool = True) -> nn.Module:
    """models.__dict__[arch](pretrained=True) of imagenet :579 with seeded random weights instead."""
    import torchvision.models as tvm

    torch.manual_seed(seed)
    model = tvm.__dict__[arch](weights=None)
    if randomize_bn:
        _randomize_bn(model, seed + 1)
    return model.eval()


@torch.no_grad()
def forward_logits(model: nn.Module, x, batch: int = 32) -> torch.Tensor:
    """`model(masked_img_tensor)` (imagenet :246) over a batch; MNIST returns the 4-tuple's last item."""
    x = torch.as_tensor(x, dtype=torch.float32)
    outs = []
    for i in range(0, x.shape[0], batch):
        o = model(x[i:i + batch])
        outs.append(o[-1] if isinstance(o, tuple) else o)
    return torch.cat(outs, 0)
