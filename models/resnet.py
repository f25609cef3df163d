"""Drop-in for the reference's models/resnet.py: `createModel(depth, data, num_classes, death_mode, death_rate)`
returns the CIFAR ResNet-(6N+2) whose evaluation forward runs in libnib.so (hand-written CUDA) instead of cuDNN.

Same constructor kwargs and the same state-dict keys as the reference (conv1, bn1, layer{1,2,3}.{i}.conv1/bn1/
conv2/bn2, fc — models/resnet.py:79-146), so `DataParallel(model).load_state_dict(checkpoint['state_dict'])`
(generate_gp_training_data_cifar.py:75,249-250) loads the shipped ResNet-56 checkpoint unchanged.  Stochastic
depth only acts in training (models/resnet.py:31); this module serves the eval path, where every block is
`relu(downsample(x) + bn2(conv2(relu(bn1(conv1(x))))))` with the parameter-free DownsampleB shortcut
(avg-pool + zero channels, :67-76)."""
from __future__ import annotations

import math

import torch.nn as nn

from ._engine_module import EngineModule


class BasicBlockWithDeathRate(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, death_rate=0.0, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.relu2 = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride
        self.death_rate = death_rate


class DownsampleB(nn.Module):
    def __init__(self, nIn, nOut, stride):
        super().__init__()
        self.avg = nn.AvgPool2d(stride)
        self.expand_ratio = nOut // nIn


class ResNetCifar(EngineModule):
    input_hw = (32, 32)

    def __init__(self, depth, death_rates=None, block=BasicBlockWithDeathRate, num_classes=10):
        assert (depth - 2) % 6 == 0, "depth should be one of 6N+2"
        super().__init__()
        n = (depth - 2) // 6
        assert death_rates is None or len(death_rates) == 3 * n
        if death_rates is None:
            death_rates = [0.0] * (3 * n)
        self.inplanes = 16
        self.conv1 = nn.Conv2d(3, 16, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(16)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = self._make_layer(block, 16, death_rates[:n])
        self.layer2 = self._make_layer(block, 32, death_rates[n:2 * n], stride=2)
        self.layer3 = self._make_layer(block, 64, death_rates[2 * n:], stride=2)
        self.avgpool = nn.AvgPool2d(8)
        self.fc = nn.Linear(64 * block.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / fan))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        self._register_engine_hooks()

    def _make_layer(self, block, planes, death_rates, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = DownsampleB(self.inplanes, planes * block.expansion, stride)
        layers = [block(self.inplanes, planes, stride, downsample=downsample, death_rate=death_rates[0])]
        self.inplanes = planes * block.expansion
        for death_rate in death_rates[1:]:
            layers.append(block(self.inplanes, planes, death_rate=death_rate))
        return nn.Sequential(*layers)


def createModel(depth, data, num_classes, death_mode="none", death_rate=0.5, **kwargs):
    assert (depth - 2) % 6 == 0, "depth should be one of 6N+2"
    print("Create ResNet-{:d} for {}".format(depth, data))
    nblocks = (depth - 2) // 2
    if death_mode == "uniform":
        death_rates = [death_rate] * nblocks
    elif death_mode == "linear":
        death_rates = [float(i + 1) * death_rate / float(nblocks) for i in range(nblocks)]
    else:
        death_rates = None
    return ResNetCifar(depth, death_rates, BasicBlockWithDeathRate, num_classes)
