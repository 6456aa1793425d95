"""NetVLAD / pyramid aggregation heads — host-side mirror of ``place_recognition/patch_aug_net/models/loupe.py``.

Parameter names, shapes and (unused) members are kept so ``state_dict`` is interchangeable with the reference:
``NetVLADBase`` still owns ``hidden1_weights`` and ``bn2`` although its forward never touches them
(loupe.py:174-188), ``MLPAttentionLayer`` still owns ``trans_conv`` / ``after_norm`` ("way 2", loupe.py:34-38).
The torch code here is the train-mode / reference-shaped path; in eval mode on CUDA,
``SpatialPyramidNetVLAD.forward`` is replaced by the fused kernels (``patchaugnet_b200.engine``).
"""
import math

import torch
import torch.nn.functional as F
from torch import nn


class MLPAttentionLayer(nn.Module):
    """x (B,C,N) -> relu(x + x * softmax_N(max_C(conv(x)))).  Reference: loupe.py:8-41."""

    def __init__(self, channels=None):
        super().__init__()
        self.mlps = nn.ModuleList(nn.Conv1d(channels[i], channels[i + 1], 1, bias=False) for i in range(len(channels) - 1))
        self.softmax = nn.Softmax(dim=-1)
        self.trans_conv = nn.Conv1d(channels[-1], channels[-1], 1)
        self.after_norm = nn.BatchNorm1d(channels[-1])
        self.act = nn.ReLU()

    def forward(self, x, return_attn=False):
        logits = x
        for mlp in self.mlps:
            logits = mlp(logits)
        weights = self.softmax(logits.max(dim=1)[0]).unsqueeze(1)   # B x 1 x N
        out = self.act(x + x * weights)
        return (out, weights) if return_attn else out


class AdaptiveFeatureAggregator(nn.Module):
    """(B, C_in, K) -> (B, C_out, 1): attention over the K columns, fc over the flattened map, BN, L2.
    Reference: loupe.py:44-66."""

    def __init__(self, C_in, K, C_out, l2_norm=True):
        super().__init__()
        self.mlpa = MLPAttentionLayer(channels=[C_in, C_in])
        self.fc = nn.Linear(C_in * K, C_out)
        self.bn = nn.BatchNorm1d(C_out)
        self.l2_norm = l2_norm

    def forward(self, x):
        x = self.mlpa(x)
        x = self.bn(self.fc(x.reshape(x.size(0), -1)))
        if self.l2_norm:
            x = F.normalize(x)
        return x.unsqueeze(-1)


class GatingContext(nn.Module):
    """x * sigmoid(BN(x @ G)).  Reference: loupe.py:330-360."""

    def __init__(self, dim, add_batch_norm=True):
        super().__init__()
        self.dim = dim
        self.add_batch_norm = add_batch_norm
        self.gating_weights = nn.Parameter(torch.randn(dim, dim) * 1 / math.sqrt(dim))
        self.sigmoid = nn.Sigmoid()
        if add_batch_norm:
            self.gating_biases = None
            self.bn1 = nn.BatchNorm1d(dim)
        else:
            self.gating_biases = nn.Parameter(torch.randn(dim) * 1 / math.sqrt(dim))
            self.bn1 = None

    def forward(self, x):
        gates = torch.matmul(x, self.gating_weights)
        gates = self.bn1(gates) if self.add_batch_norm else gates + self.gating_biases
        return x * self.sigmoid(gates)


class NetVLADBase(nn.Module):
    """(B, C, N, 1) -> (B, C, K) intra-normalised VLAD.  Reference: loupe.py:159-222."""

    def __init__(self, feature_size, max_samples, cluster_size, output_dim, gating=True, add_batch_norm=True):
        super().__init__()
        self.feature_size = feature_size
        self.max_samples = max_samples
        self.output_dim = output_dim
        self.gating = gating
        self.add_batch_norm = add_batch_norm
        self.cluster_size = cluster_size
        self.softmax = nn.Softmax(dim=-1)
        s = 1 / math.sqrt(feature_size)
        self.cluster_weights = nn.Parameter(torch.randn(feature_size, cluster_size) * s)
        self.cluster_weights2 = nn.Parameter(torch.randn(1, feature_size, cluster_size) * s)
        self.hidden1_weights = nn.Parameter(torch.randn(feature_size * cluster_size, output_dim) * s)
        if add_batch_norm:
            self.cluster_biases = None
            self.bn1 = nn.BatchNorm1d(cluster_size)
        else:
            self.cluster_biases = nn.Parameter(torch.randn(cluster_size) * s)
            self.bn1 = None
        self.bn2 = nn.BatchNorm1d(output_dim)
        if gating:
            self.context_gating = GatingContext(output_dim, add_batch_norm=add_batch_norm)

    def forward(self, x):
        x = x.transpose(1, 3).contiguous().view(-1, self.max_samples, self.feature_size)       # B x N x C
        act = torch.matmul(x, self.cluster_weights)                                             # B x N x K
        if self.add_batch_norm:
            act = self.bn1(act.view(-1, self.cluster_size)).view(-1, self.max_samples, self.cluster_size)
        else:
            act = act + self.cluster_biases
        act = self.softmax(act)
        a = act.sum(-2, keepdim=True) * self.cluster_weights2                                   # B x C x K
        vlad = torch.matmul(act.transpose(2, 1), x).transpose(2, 1) - a                         # B x C x K
        return F.normalize(vlad, dim=1, p=2).contiguous()


class SpatialPyramidNetVLAD(nn.Module):
    """One NetVLAD per pyramid level, then one of six aggregation variants.  Reference: loupe.py:225-329."""

    def __init__(self, feature_size, max_samples, cluster_size, output_dim, gating=True, aggregation_type=False,
                 add_batch_norm=True):
        super().__init__()
        assert len(feature_size) == len(max_samples) == len(cluster_size) == len(output_dim)
        nl = len(feature_size)
        self.vlads = nn.ModuleList(
            NetVLADBase(feature_size[i], max_samples[i], cluster_size[i], output_dim[i], gating, add_batch_norm)
            for i in range(nl))
        sum_k = sum(cluster_size)
        self.gating = gating
        if gating:
            self.context_gating = GatingContext(output_dim[0], add_batch_norm=add_batch_norm)
        self.aggregation_type = aggregation_type
        s = 1 / math.sqrt(feature_size[0])
        if aggregation_type == 0:
            self.hidden_weights = nn.Parameter(torch.randn(feature_size[0] * sum_k, output_dim[0]) * s)
            self.bn = nn.BatchNorm1d(output_dim[0])
        elif aggregation_type == 1:
            self.afa_scales = nn.ModuleList(
                AdaptiveFeatureAggregator(output_dim[i], cluster_size[i], output_dim[i]) for i in range(nl))
            self.afa = AdaptiveFeatureAggregator(output_dim[0], nl, output_dim[0])
        elif aggregation_type == 2:
            self.afa = AdaptiveFeatureAggregator(output_dim[0], sum_k, output_dim[0])
        elif aggregation_type == 4:
            self.afa_scales = nn.ModuleList(
                AdaptiveFeatureAggregator(output_dim[i], cluster_size[i], output_dim[i]) for i in range(nl))
            self.hidden_weights = nn.Parameter(torch.randn(feature_size[0] * nl, output_dim[0]) * s)
            self.bn = nn.BatchNorm1d(output_dim[0])
        elif aggregation_type == 5:
            self.hidden_weights = nn.ParameterList()
            self.bns = nn.ModuleList()
            for i in range(nl):
                self.hidden_weights.append(nn.Parameter(
                    torch.randn(feature_size[i] * cluster_size[i], output_dim[i]) * 1 / math.sqrt(feature_size[i])))
                self.bns.append(nn.BatchNorm1d(output_dim[i]))
            self.afa = AdaptiveFeatureAggregator(output_dim[0], nl, output_dim[0])

    def forward(self, features=None):
        return self.aggregate([vlad(f) for vlad, f in zip(self.vlads, features or [])])

    def aggregate(self, v):
        """The aggregation variant over the per-level VLADs v[i] (B, C, K_i) (reference loupe.py:289-328).  Also the tail of
        the fused engine for every variant but the configured one (type 2 without gating has its own kernels)."""
        t = self.aggregation_type
        if t == 0:
            cat = torch.cat(v, dim=-1)
            out = F.normalize(self.bn(torch.matmul(cat.view(cat.size(0), -1), self.hidden_weights)))
        elif t == 1:
            out = self.afa(torch.cat([a(x) for a, x in zip(self.afa_scales, v)], dim=-1)).squeeze(-1)
        elif t == 2:
            out = self.afa(torch.cat(v, dim=-1)).squeeze(-1)
        elif t == 3:
            cat = torch.cat(v, dim=-1)
            out = F.normalize(F.max_pool2d(cat, kernel_size=[1, cat.size(2)]).squeeze(-1))
        elif t == 4:
            cat = torch.cat([a(x) for a, x in zip(self.afa_scales, v)], dim=-1)
            out = F.normalize(self.bn(torch.matmul(cat.view(cat.size(0), -1), self.hidden_weights)))
        elif t == 5:
            per = []
            for i, x in enumerate(v):
                x = self.bns[i](torch.matmul(x.view(x.size(0), -1), self.hidden_weights[i]))
                per.append(F.normalize(x).unsqueeze(-1))
            out = self.afa(torch.cat(per, dim=-1))
        else:
            raise ValueError(f"unknown aggregation_type {t}")
        if self.gating:
            out = self.context_gating(out)
        return out
