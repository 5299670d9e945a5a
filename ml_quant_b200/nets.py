"""Callers of the hot path: the two networks the reference's examples build.

Own implementations with the constructor arguments and state_dict key layout of
quant/models/resnet.py (RegularBasicBlock :28-97, XnorBasicBlock :100-190, QResNet :193-397) and
quant/models/lenet.py (QLeNet5 :21-94), so the shipped YAML configs and checkpoints apply unchanged.
They exist because the benchmark and the parity tests need the reference's networks on a box that
does not have the reference; everything quantized inside them is ``QuantConv2d``.
"""
from typing import Callable, Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .binary.binary_conv import QuantConv2d

non_linearity_map = {'relu': nn.ReLU, 'prelu': nn.PReLU, 'identity': nn.Identity}


def _fp_shortcut(in_planes: int, planes: int, stride: int, bias: bool) -> nn.Sequential:
    if stride == 1 and in_planes == planes:
        return nn.Sequential()
    return nn.Sequential(nn.Conv2d(in_planes, planes, kernel_size=1, stride=stride, bias=bias),
                         nn.BatchNorm2d(planes))


def _two(nonlins: List[str]) -> List[str]:
    if len(nonlins) != 2:
        raise ValueError('There should be 2 non-linearities.')
    return nonlins


class RegularBasicBlock(nn.Module):
    """conv-bn-nonlin, conv-bn, + shortcut, nonlin (full-precision downsampling shortcut)."""

    def __init__(self, in_planes: int, planes: int, x_quant: str, w_quant: str, nonlins: List[str],
                 stride: int = 1, clamp: Optional[Dict] = None, moving_average_mode: str = 'off',
                 moving_average_momentum: float = 0.99) -> None:
        super().__init__()
        n1, n2 = _two(nonlins)
        q = dict(clamp=clamp, moving_average_mode=moving_average_mode,
                 moving_average_momentum=moving_average_momentum, padding=1, bias=False)
        self.conv1 = QuantConv2d(x_quant, w_quant, in_planes, planes, 3, stride=stride, **q)
        self.bn1 = nn.BatchNorm2d(planes)
        self.nonlin1 = non_linearity_map[n1]()
        self.conv2 = QuantConv2d(x_quant, w_quant, planes, planes, 3, stride=1, **q)
        self.bn2 = nn.BatchNorm2d(planes)
        self.nonlin2 = non_linearity_map[n2]()
        self.shortcut = _fp_shortcut(in_planes, planes, stride, bias=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        y = self.nonlin1(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y)) + self.shortcut(x)
        return self.nonlin2(y)


class XnorBasicBlock(nn.Module):
    """bn-quantconv-nonlin twice (XNOR-Net ordering), optionally with Bi-Real double shortcuts."""

    def __init__(self, in_planes: int, planes: int, x_quant: str, w_quant: str, nonlins: List[str],
                 stride: int = 1, double_shortcut: bool = False, clamp: Optional[Dict] = None,
                 moving_average_mode: str = 'off', moving_average_momentum: float = 0.99) -> None:
        super().__init__()
        n1, n2 = _two(nonlins)
        self.double_shortcut = double_shortcut
        q = dict(clamp=clamp, moving_average_mode=moving_average_mode,
                 moving_average_momentum=moving_average_momentum, padding=1, bias=True)
        self.bn1 = nn.BatchNorm2d(in_planes)
        self.conv1 = QuantConv2d(x_quant, w_quant, in_planes, planes, 3, stride=stride, **q)
        self.nonlin1 = non_linearity_map[n1]()
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv2 = QuantConv2d(x_quant, w_quant, planes, planes, 3, stride=1, **q)
        self.nonlin2 = non_linearity_map[n2]()
        self.shortcut = _fp_shortcut(in_planes, planes, stride, bias=True)

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        first = self.nonlin1(self.conv1(self.bn1(x)))
        if self.double_shortcut:
            first = first + self.shortcut(x)
            return self.nonlin2(self.conv2(self.bn2(first))) + first
        return self.nonlin2(self.conv2(self.bn2(first)) + self.shortcut(x))


class QResNet(nn.Module):
    """ResNet of basic blocks: fp stem (layer0), 3-4 quantized stages, fp classifier."""

    def __init__(self, loss_fn: Callable[..., torch.Tensor], block: str, layer0: dict, layer1: dict,
                 layer2: dict, layer3: dict, layer4: Optional[dict], nonlins: List[str], num_blocks: List[int],
                 output_classes: int, moving_average_mode: str = 'off',
                 moving_average_momentum: float = 0.99) -> None:
        super().__init__()
        setattr(self, 'loss_fn', loss_fn)
        kinds = {'regular': RegularBasicBlock, 'xnor': XnorBasicBlock}
        if block not in kinds:
            raise ValueError(f'Block {block} is not supported.')
        width = layer0['n_in_channels']
        self.conv1 = nn.Conv2d(3, width, kernel_size=layer0['kernel_size'], stride=layer0['stride'],
                               padding=layer0['padding'], bias=layer0['bias'])
        pool = layer0['maxpool']
        if pool['type'] == 'identity':
            self.maxpool: nn.Module = nn.Identity()
        elif pool['type'] == 'maxpool2d':
            self.maxpool = nn.MaxPool2d(kernel_size=pool['kernel_size'], stride=pool['stride'],
                                        padding=pool['padding'])
        else:
            raise ValueError(f"maxpool type {pool['type']} is not supported.")
        self.bn1 = nn.BatchNorm2d(width)
        # the stem modules are registered twice (attribute and blocks[0]) exactly like the reference,
        # so state_dict carries both key sets and checkpoints are interchangeable
        self.blocks = nn.ModuleList([nn.Sequential(self.conv1, self.bn1, nn.ReLU(inplace=True), self.maxpool)])
        planes = width
        for i, (cfg, count) in enumerate(zip([layer1, layer2, layer3, layer4], num_blocks)):
            if cfg is None:
                continue
            out_planes = width << i
            for j in range(count):
                self.blocks.append(kinds[block](
                    planes, out_planes, nonlins=nonlins, stride=(2 if (i > 0 and j == 0) else 1),
                    moving_average_mode=moving_average_mode, moving_average_momentum=moving_average_momentum,
                    **cfg))
                planes = out_planes
        self.linear_classifier = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten(),
                                               nn.Linear(planes, output_classes))

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        for blk in self.blocks:
            x = blk(x)
        return self.linear_classifier(x)


class QLeNet5(nn.Module):
    """LeNet-5: fp conv1, one QuantConv2d (conv2), fp classifier."""

    def __init__(self, loss_fn: Callable[..., torch.Tensor], conv1_filters: int = 20, conv2_filters: int = 50,
                 output_classes: int = 10, x_quant: str = 'fp', w_quant: str = 'fp', clamp: Optional[Dict] = None,
                 moving_average_mode: str = 'off', moving_average_momentum: float = 0.99) -> None:
        super().__init__()
        setattr(self, 'loss_fn', loss_fn)
        self.conv1_filters, self.conv2_filters, self.output_classes = conv1_filters, conv2_filters, output_classes
        self.x_quant, self.w_quant = x_quant, w_quant
        self.conv1 = nn.Conv2d(1, conv1_filters, 5, stride=1)
        self.bn_conv1 = nn.BatchNorm2d(conv1_filters, eps=1e-4, momentum=0.1, affine=False)
        self.conv2 = QuantConv2d(x_quant, w_quant, conv1_filters, conv2_filters, 5, clamp,
                                 moving_average_mode, moving_average_momentum, stride=1)
        self.bn_conv2 = nn.BatchNorm2d(conv1_filters, eps=1e-4, momentum=0.1, affine=False)
        self.fc1 = nn.Linear(conv2_filters * 16, conv2_filters * output_classes)
        self.fc2 = nn.Linear(conv2_filters * output_classes, output_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        x = F.max_pool2d(self.bn_conv1(F.relu(self.conv1(x), inplace=True)), kernel_size=2, stride=2)
        x = F.max_pool2d(F.relu(self.conv2(self.bn_conv2(x)), inplace=True), kernel_size=2, stride=2)
        x = F.relu(self.fc1(x.view(-1, self.conv2_filters * 16)), inplace=True)
        return F.log_softmax(self.fc2(x), dim=1)
