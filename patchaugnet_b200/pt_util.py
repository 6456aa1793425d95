"""SharedMLP — host-side mirror of the reference's ``utils/model_util/pt_util.py:16-41, 98-151, 188-219``.

Only the part the hot path uses: a stack of ``Conv2d(1x1, bias=False) -> BatchNorm2d -> ReLU`` blocks whose
``state_dict`` keys are byte-identical to the reference's (``layer{i}.conv.weight``, ``layer{i}.bn.bn.weight`` ...),
so reference checkpoints load unchanged.  In eval mode the fused CUDA engine consumes these parameters folded
(``patchaugnet_b200.engine``); the torch forward below is the train-mode path (batch statistics).
"""
from typing import List

import torch
import torch.nn as nn

from . import _lib as L

FUSED_TRAIN_BN_RELU = True       # train mode on CUDA: BatchNorm (batch statistics) + ReLU as one fused kernel pair (csrc/bn_train.cu)


class _BnReluTrain(torch.autograd.Function):
    """y = relu(batch_norm_train(x)) for x (B, C, ...) on hand-written kernels; updates the running statistics like nn.BatchNorm."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps):
        x = x.contiguous()
        B, C = x.shape[0], x.shape[1]
        S = x.numel() // (B * C)
        y = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        invstd = torch.empty_like(mean)
        ws = torch.empty(L.lib().pab_bn_train_workspace_bytes(C), dtype=torch.uint8, device=x.device)
        L.check(L.lib().pab_bn_relu_train_forward(B, C, S, L.ptr(x), L.ptr(weight), L.ptr(bias), float(eps), float(momentum),
                                                  L.ptr(running_mean), L.ptr(running_var), L.ptr(mean), L.ptr(invstd), L.ptr(y),
                                                  L.ptr(ws), L.stream_ptr()), "bn_relu_train_forward")
        ctx.save_for_backward(x, y, weight, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, weight, mean, invstd = ctx.saved_tensors
        dy = dy.contiguous()
        B, C = x.shape[0], x.shape[1]
        S = x.numel() // (B * C)
        dx = torch.empty_like(x)
        dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
        dbeta = torch.empty_like(dgamma)
        ws = torch.empty(L.lib().pab_bn_train_workspace_bytes(C), dtype=torch.uint8, device=x.device)
        L.check(L.lib().pab_bn_relu_train_backward(B, C, S, L.ptr(dy), L.ptr(y), L.ptr(x), L.ptr(weight), L.ptr(mean), L.ptr(invstd),
                                                   L.ptr(dx), L.ptr(dgamma), L.ptr(dbeta), L.ptr(ws), L.stream_ptr()),
                "bn_relu_train_backward")
        return dx, dgamma, dbeta, None, None, None, None


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

    def forward(self, x):
        bn = self.bn.bn if hasattr(self, "bn") else None
        if (FUSED_TRAIN_BN_RELU and self.training and bn is not None and hasattr(self, "activation") and x.is_cuda
                and x.dtype == torch.float32 and bn.track_running_stats and bn.momentum is not None and bn.affine):
            x = self.conv(x)
            if bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
            return _BnReluTrain.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps)
        return super().forward(x)


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
