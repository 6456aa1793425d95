"""PPT-Net aggregation head — mirror of ``place_recognition/pptnet_origin/models/loupe.py``.

Differs from the PatchAugNet head (``patchaugnet_b200/loupe.py``): each ``NetVLADBase`` returns its VLAD flattened to
(B, C*K) (pptnet_origin/models/loupe.py:69-70), the four levels are named ``vlad0..vlad3``, and the fusion is a dense
``hidden_weights`` (21760 x 256) + ``bn2`` + context gating (:73-105).  Unused members (``hidden1_weights``, ``bn2`` and
``context_gating`` inside each level) are kept so ``state_dict`` matches the reference.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .loupe import GatingContext


class NetVLADBase(nn.Module):
    """(B, C, N, 1) -> (B, C*K).  Reference: pptnet_origin/models/loupe.py:6-71."""

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
        x = x.transpose(1, 3).contiguous().view(-1, self.max_samples, self.feature_size)
        act = torch.matmul(x, self.cluster_weights)
        if self.add_batch_norm:
            act = self.bn1(act.view(-1, self.cluster_size)).view(-1, self.max_samples, self.cluster_size)
        else:
            act = act + self.cluster_biases
        act = self.softmax(act)
        a = act.sum(-2, keepdim=True) * self.cluster_weights2
        vlad = torch.matmul(act.transpose(2, 1), x).transpose(2, 1) - a
        vlad = F.normalize(vlad, dim=1, p=2).contiguous()
        return vlad.view(-1, self.cluster_size * self.feature_size)


class SpatialPyramidNetVLAD(nn.Module):
    """Four NetVLAD levels -> hidden_weights -> BN -> context gating.  Reference: pptnet_origin/models/loupe.py:73-105."""

    def __init__(self, feature_size, max_samples, cluster_size, output_dim, gating=True, add_batch_norm=True):
        super().__init__()
        for i in range(4):
            setattr(self, f"vlad{i}", NetVLADBase(feature_size[i], max_samples[i], cluster_size[i], output_dim[i], gating,
                                                   add_batch_norm))
        self.hidden_weights = nn.Parameter(
            torch.randn(feature_size[0] * sum(cluster_size[:4]), output_dim[0]) * 1 / math.sqrt(feature_size[0]))
        self.bn2 = nn.BatchNorm1d(output_dim[0])
        self.gating = gating
        if gating:
            self.context_gating = GatingContext(output_dim[0], add_batch_norm=add_batch_norm)

    def forward(self, f0, f1, f2, f3):
        vlad = torch.cat((self.vlad0(f0), self.vlad1(f1), self.vlad2(f2), self.vlad3(f3)), dim=-1)
        vlad = self.bn2(torch.matmul(vlad, self.hidden_weights))
        return self.context_gating(vlad) if self.gating else vlad
