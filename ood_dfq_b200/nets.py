"""Shape-faithful carrier networks for the benchmark configs (plain torch.nn, no kernels here).

Convolutions stay on PyTorch/cuDNN (BASELINE.json north_star); these definitions only
give the quantisation path the tensors it sees in the reference:

* ``resnet20_cifar`` / ``resnet18_imagenet`` follow the module tree of the un-vendored
  ``pytorchcv`` models the reference loads with ``ptcv_get_model`` (main_direct.py:380-397),
  as recoverable from its checkpoint key map (main_direct.py:253-294):
  ``features.init_block.conv.{conv,bn}``, ``features.stage{i}.unit{j}.body.conv{1,2}.{conv,bn}``,
  ``features.stage{i}.unit1.identity_conv.{conv,bn}``, ``output``.  Every conv block and every
  unit owns its OWN ReLU module, which is what makes ``quantize_model`` produce 17 (ResNet-18)
  / 19 (ResNet-20) ``QuantAct`` sites.
* ``resnet18_small`` mirrors the 28x28 MedMNIST ResNet-18 of the reference's models.py:9-111
  (same attribute names, so its checkpoints load).
"""
from __future__ import annotations

import torch
from torch import nn


class ConvBlock(nn.Module):
    """conv -> bn -> (relu): attributes ``conv``, ``bn``, ``activ`` as in pytorchcv's ConvBlock."""

    def __init__(self, cin, cout, kernel, stride=1, padding=0, activate=True):
        super().__init__()
        self.activate = activate
        self.conv = nn.Conv2d(cin, cout, kernel, stride=stride, padding=padding, bias=False)
        self.bn = nn.BatchNorm2d(cout)
        if activate:
            self.activ = nn.ReLU(inplace=True)

    def forward(self, x):
        x = self.bn(self.conv(x))
        return self.activ(x) if self.activate else x


class ResBody(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = ConvBlock(cin, cout, 3, stride=stride, padding=1)
        self.conv2 = ConvBlock(cout, cout, 3, padding=1, activate=False)

    def forward(self, x):
        return self.conv2(self.conv1(x))


class ResUnit(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.resize_identity = (cin != cout) or (stride != 1)
        self.body = ResBody(cin, cout, stride)
        if self.resize_identity:
            self.identity_conv = ConvBlock(cin, cout, 1, stride=stride, activate=False)
        self.activ = nn.ReLU(inplace=True)

    def forward(self, x):
        identity = self.identity_conv(x) if self.resize_identity else x
        return self.activ(self.body(x) + identity)


class InitBlockImageNet(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = ConvBlock(cin, cout, 7, stride=2, padding=3)
        self.pool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)

    def forward(self, x):
        return self.pool(self.conv(x))


class PtcvResNet(nn.Module):
    def __init__(self, channels, init_channels, num_classes, imagenet, in_channels=3, final_pool=7):
        super().__init__()
        self.features = nn.Sequential()
        if imagenet:
            self.features.add_module("init_block", InitBlockImageNet(in_channels, init_channels))
        else:
            self.features.add_module("init_block", ConvBlock(in_channels, init_channels, 3, padding=1))
        cin = init_channels
        for i, stage_channels in enumerate(channels):
            stage = nn.Sequential()
            for j, cout in enumerate(stage_channels):
                stride = 2 if (j == 0 and i != 0) else 1
                stage.add_module(f"unit{j + 1}", ResUnit(cin, cout, stride))
                cin = cout
            self.features.add_module(f"stage{i + 1}", stage)
        self.features.add_module("final_pool", nn.AvgPool2d(kernel_size=final_pool, stride=1))
        self.output = nn.Linear(cin, num_classes)

    def forward(self, x):
        x = self.features(x)
        return self.output(x.view(x.size(0), -1))


def resnet18_imagenet(num_classes=1000):
    """pytorchcv ``resnet18`` shape: 224x224 input, 20 convs, 20 BNs, 17 ReLU sites."""
    return PtcvResNet([[64, 64], [128, 128], [256, 256], [512, 512]], 64, num_classes, imagenet=True, final_pool=7)


def resnet20_cifar(num_classes=10):
    """pytorchcv ``resnet20_cifar10/100`` shape: 32x32 input, 21 convs, 21 BNs, 19 ReLU sites."""
    return PtcvResNet([[16] * 3, [32] * 3, [64] * 3], 16, num_classes, imagenet=False, final_pool=8)


# ------------------------------------------------------------------------------ 28x28 ResNet-18
class SmallBlock(nn.Module):
    """Two 3x3 convs with a projection shortcut when the shape changes; two separate ReLU modules."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.shortcut = nn.Sequential()
        if stride != 1 or cin != cout:
            self.shortcut = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), nn.BatchNorm2d(cout))
        self.relu2 = nn.ReLU(inplace=True)

    def forward(self, x):
        y = self.relu1(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        y = y + self.shortcut(x)
        return self.relu2(y)


class SmallResNet18(nn.Module):
    def __init__(self, in_channels, num_classes, img_size=28):
        super().__init__()
        self.img_size = img_size
        self.conv1 = nn.Conv2d(in_channels, 64, 3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu1 = nn.ReLU(inplace=True)
        widths, cin = (64, 128, 256, 512), 64
        for i, w in enumerate(widths):
            blocks = [SmallBlock(cin, w, 1 if i == 0 else 2), SmallBlock(w, w, 1)]
            setattr(self, f"layer{i + 1}", nn.Sequential(*blocks))
            cin = w
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.linear = nn.Linear(512, num_classes)

    def forward(self, x):
        x = self.relu1(self.bn1(self.conv1(x)))
        for i in range(4):
            x = getattr(self, f"layer{i + 1}")(x)
        x = self.avgpool(x)
        return self.linear(x.view(x.size(0), -1))


def resnet18_small(in_channels=3, num_classes=9, img_size=28):
    """The reference's own 28x28 ResNet-18 (models.py:110-111) for the MedMNIST configs."""
    return SmallResNet18(in_channels, num_classes, img_size)


def perturb_bn_stats(model: nn.Module, seed=2):
    """Give every BN non-trivial running statistics (rm ~ N(0, 0.1^2), rv ~ U(0.5, 1.5)) -- SURVEY 8(d) config 1."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.modules.batchnorm._BatchNorm):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
    return model
