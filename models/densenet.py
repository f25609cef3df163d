"""Drop-in for the reference's models/densenet.py: `createModel(data, depth, growth_rate, num_classes, drop_rate,
num_init_features, compression, bn_size)` returns the CIFAR DenseNet(-BC) (models/densenet.py:44-120) whose
evaluation forward runs in libnib.so.

The reference registers sub-modules as 'norm.1', 'conv.1', ... (models/densenet.py:16-27), which raises KeyError
on torch >= 1.0; the legal names norm1/conv1/norm2/conv2 are used here — no DenseNet
checkpoint is shipped, so no key compatibility is lost.  `data='imagenet'` in the reference is dimensionally
inconsistent (3 blocks, no stem stride); ImageNet DenseNets go through torchvision.models.densenet121
(`-a densenet121`, generate_gp_training_data_imagenet.py:579), which Classifier.from_torch lowers directly."""
from __future__ import annotations

import torch.nn as nn

from ._engine_module import EngineModule


class _DenseLayer(nn.Module):
    def __init__(self, num_input_features, growth_rate, bn_size, drop_rate):
        super().__init__()
        if bn_size > 0:   # bottleneck (BC)
            self.norm1 = nn.BatchNorm2d(num_input_features)
            self.relu1 = nn.ReLU(inplace=True)
            self.conv1 = nn.Conv2d(num_input_features, bn_size * growth_rate, kernel_size=1, stride=1, bias=False)
            self.norm2 = nn.BatchNorm2d(bn_size * growth_rate)
            self.relu2 = nn.ReLU(inplace=True)
            self.conv2 = nn.Conv2d(bn_size * growth_rate, growth_rate, kernel_size=3, stride=1, padding=1, bias=False)
        else:
            self.norm1 = nn.BatchNorm2d(num_input_features)
            self.relu1 = nn.ReLU(inplace=True)
            self.conv1 = nn.Conv2d(num_input_features, growth_rate, kernel_size=3, stride=1, padding=1, bias=False)
        self.drop_rate = drop_rate   # dropout acts in training only


class DenseNet(EngineModule):
    input_hw = (32, 32)

    def __init__(self, growth_rate=12, block_config=(16, 16, 16), compression=0.5, num_init_features=24, bn_size=4,
                 drop_rate=0, avgpool_size=8, num_classes=10):
        super().__init__()
        assert 0 < compression <= 1, "compression of densenet should be between 0 and 1"
        self.avgpool_size = avgpool_size
        self.stem = nn.Conv2d(3, num_init_features, kernel_size=3, stride=1, padding=1, bias=False)
        nf = num_init_features
        blocks = []
        for i, num_layers in enumerate(block_config):
            layers = []
            for _ in range(num_layers):
                layers.append(_DenseLayer(nf, growth_rate, bn_size, drop_rate))
                nf += growth_rate
            blocks.append(nn.Sequential(*layers))
            if i != len(block_config) - 1:
                no = int(nf * compression)
                blocks.append(nn.Sequential(nn.BatchNorm2d(nf), nn.ReLU(inplace=True),
                                            nn.Conv2d(nf, no, kernel_size=1, stride=1, bias=False),
                                            nn.AvgPool2d(kernel_size=2, stride=2)))
                nf = no
        self.blocks = nn.Sequential(*blocks)
        self.norm5 = nn.BatchNorm2d(nf)
        self.classifier = nn.Linear(nf, num_classes)
        self._register_engine_hooks()


def createModel(data, depth=100, growth_rate=12, num_classes=10, drop_rate=0, num_init_features=24, compression=0.5,
                bn_size=4, **kwargs):
    if data not in ("cifar10", "cifar100"):
        raise NotImplementedError("models/densenet.py supports the CIFAR geometry; for ImageNet use torchvision "
                                  "densenet121 (generate_gp_training_data_imagenet.py -a densenet121)")
    n = (depth - 4) // 3
    if bn_size > 0:
        n //= 2
    print("Create DenseNet{}-{:d} for {}".format("-BC" if bn_size > 0 else "", depth, data))
    return DenseNet(growth_rate=growth_rate, block_config=(n, n, n), compression=compression,
                    num_init_features=num_init_features, bn_size=bn_size, drop_rate=drop_rate, avgpool_size=8,
                    num_classes=num_classes)
