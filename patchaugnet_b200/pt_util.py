"""SharedMLP — host-side mirror of the reference's ``utils/model_util/pt_util.py:16-41, 98-151, 188-219``.

Only the part the hot path uses: a stack of ``Conv2d(1x1, bias=False) -> BatchNorm2d -> ReLU`` blocks whose
``state_dict`` keys are byte-identical to the reference's (``layer{i}.conv.weight``, ``layer{i}.bn.bn.weight`` ...),
so reference checkpoints load unchanged.  In eval mode the fused CUDA engine consumes these parameters folded
(``patchaugnet_b200.engine``); the torch forward below is the train-mode path (batch statistics).
"""
from typing import List

import torch.nn as nn


class _BN(nn.Sequential):
    """``bn.bn``: the reference wraps the norm in a one-element Sequential (pt_util.py:70-95)."""

    def __init__(self, channels: int, norm=nn.BatchNorm2d):
        super().__init__()
        self.add_module("bn", norm(channels))
        nn.init.constant_(self[0].weight, 1.0)
        nn.init.constant_(self[0].bias, 0)


class _ConvBlock(nn.Sequential):
    """conv -> bn -> activation with the reference's sub-module names (pt_util.py:98-151, post-activation order)."""

    def __init__(self, c_in, c_out, conv, norm, bn=True, activation=True, bias=True):
        super().__init__()
        unit = conv(c_in, c_out, kernel_size=1, bias=bias and not bn)
        nn.init.kaiming_normal_(unit.weight)
        if unit.bias is not None:
            nn.init.constant_(unit.bias, 0)
        self.add_module("conv", unit)
        if bn:
            self.add_module("bn", _BN(c_out, norm))
        if activation:
            self.add_module("activation", nn.ReLU(inplace=True))


class Conv2d(_ConvBlock):
    def __init__(self, c_in, c_out, bn=False, activation=True, bias=True):
        super().__init__(c_in, c_out, nn.Conv2d, nn.BatchNorm2d, bn=bn, activation=activation, bias=bias)


class Conv1d(_ConvBlock):
    def __init__(self, c_in, c_out, bn=False, activation=True, bias=True):
        super().__init__(c_in, c_out, nn.Conv1d, nn.BatchNorm1d, bn=bn, activation=activation, bias=bias)


class SharedMLP(nn.Sequential):
    """Point-wise MLP as 1x1 convolutions over (B, C, M, K).  Reference: pt_util.py:16-41."""

    def __init__(self, args: List[int], *, bn: bool = False, name: str = ""):
        super().__init__()
        self.spec = list(args)
        for i in range(len(args) - 1):
            self.add_module(f"{name}layer{i}", Conv2d(args[i], args[i + 1], bn=bn))


class SharedMLP_1d(nn.Sequential):
    """Reference: pt_util.py:43-68."""

    def __init__(self, args: List[int], *, bn: bool = False, name: str = ""):
        super().__init__()
        self.spec = list(args)
        for i in range(len(args) - 1):
            self.add_module(f"{name}layer{i}", Conv1d(args[i], args[i + 1], bn=bn))
