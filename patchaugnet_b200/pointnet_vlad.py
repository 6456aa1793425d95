"""PointNetVLAD — mirror of ``place_recognition/pointnet_vlad/PointNetVlad.py`` (BASELINE.json configs[0]: the
reference's own CPU-runnable case; pure PyTorch in the reference, and kept pure PyTorch here — it has no custom kernel
to replace).  Same module / parameter names (``point_net.stn.conv1`` ... ``net_vlad.hidden1_weights``), so reference
checkpoints load unchanged (78 state_dict entries, 19,779,145 parameters).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .loupe import GatingContext


class NetVLADLoupe(nn.Module):
    """(B, C, N, 1) -> (B, output_dim): VLAD, intra + global L2, hidden fc, BN, gating.  Reference: PointNetVlad.py:12-81."""

    def __init__(self, feature_size, max_samples, cluster_size, output_dim, gating=True, add_batch_norm=True, is_training=True):
        super().__init__()
        self.feature_size = feature_size
        self.max_samples = max_samples
        self.output_dim = output_dim
        self.is_training = is_training
        self.gating = gating
        self.add_batch_norm = add_batch_norm
        self.cluster_size = cluster_size
        self.softmax = nn.Softmax(dim=-1)
        s = 1 / math.sqrt(feature_size)
        self.cluster_weights = nn.Parameter(torch.randn(feature_size, cluster_size) * s)
        self.cluster_weights2 = nn.Parameter(torch.randn(1, feature_size, cluster_size) * s)
        self.hidden1_weights = nn.Parameter(torch.randn(cluster_size * feature_size, output_dim) * s)
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
        vlad = F.normalize(vlad, dim=1, p=2).contiguous().view(-1, self.cluster_size * self.feature_size)
        vlad = F.normalize(vlad, dim=1, p=2)
        vlad = self.bn2(torch.matmul(vlad, self.hidden1_weights))
        return self.context_gating(vlad) if self.gating else vlad


class STN3d(nn.Module):
    """Spatial transformer predicting a k x k matrix (initialised to identity).  Reference: PointNetVlad.py:124-180."""

    def __init__(self, num_points=2500, k=3, use_bn=True):
        super().__init__()
        self.k = k
        self.kernel_size = 3 if k == 3 else 1
        self.channels = 1 if k == 3 else k
        self.num_points = num_points
        self.use_bn = use_bn
        self.conv1 = nn.Conv2d(self.channels, 64, (1, self.kernel_size))
        self.conv2 = nn.Conv2d(64, 128, (1, 1))
        self.conv3 = nn.Conv2d(128, 1024, (1, 1))
        self.mp1 = nn.MaxPool2d((num_points, 1), 1)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, k * k)
        self.fc3.weight.data.zero_()
        self.fc3.bias.data.zero_()
        self.relu = nn.ReLU()
        if use_bn:
            self.bn1 = nn.BatchNorm2d(64)
            self.bn2 = nn.BatchNorm2d(128)
            self.bn3 = nn.BatchNorm2d(1024)
            self.bn4 = nn.BatchNorm1d(512)
            self.bn5 = nn.BatchNorm1d(256)

    def forward(self, x):
        b = x.size(0)
        n = (lambda m, y: m(y)) if self.use_bn else (lambda m, y: y)
        x = F.relu(n(getattr(self, "bn1", None), self.conv1(x)))
        x = F.relu(n(getattr(self, "bn2", None), self.conv2(x)))
        x = F.relu(n(getattr(self, "bn3", None), self.conv3(x)))
        x = self.mp1(x).view(-1, 1024)
        x = F.relu(n(getattr(self, "bn4", None), self.fc1(x)))
        x = F.relu(n(getattr(self, "bn5", None), self.fc2(x)))
        x = self.fc3(x) + torch.eye(self.k, dtype=x.dtype, device=x.device).view(1, self.k * self.k).repeat(b, 1)
        return x.view(-1, self.k, self.k)


class PointNetfeat(nn.Module):
    """Input STN, 5 point-wise convs, optional 64-D feature STN.  Reference: PointNetVlad.py:183-232."""

    def __init__(self, num_points=2500, global_feat=True, feature_transform=False, max_pool=True):
        super().__init__()
        self.stn = STN3d(num_points=num_points, k=3, use_bn=False)
        self.feature_trans = STN3d(num_points=num_points, k=64, use_bn=False)
        self.apply_feature_trans = feature_transform
        self.conv1 = nn.Conv2d(1, 64, (1, 3))
        self.conv2 = nn.Conv2d(64, 64, (1, 1))
        self.conv3 = nn.Conv2d(64, 64, (1, 1))
        self.conv4 = nn.Conv2d(64, 128, (1, 1))
        self.conv5 = nn.Conv2d(128, 1024, (1, 1))
        self.bn1 = nn.BatchNorm2d(64)
        self.bn2 = nn.BatchNorm2d(64)
        self.bn3 = nn.BatchNorm2d(64)
        self.bn4 = nn.BatchNorm2d(128)
        self.bn5 = nn.BatchNorm2d(1024)
        self.mp1 = nn.MaxPool2d((num_points, 1), 1)
        self.num_points = num_points
        self.global_feat = global_feat
        self.max_pool = max_pool

    def forward(self, x):
        b = x.size(0)
        trans = self.stn(x)
        x = torch.matmul(x.reshape(b, -1, 3), trans).view(b, 1, -1, 3)
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.relu(self.bn2(self.conv2(x)))
        pointfeat = x
        if self.apply_feature_trans:
            f_trans = self.feature_trans(x)
            x = torch.matmul(x.reshape(b, 64, -1).transpose(1, 2), f_trans).transpose(1, 2).contiguous().view(b, 64, -1, 1)
        x = F.relu(self.bn3(self.conv3(x)))
        x = F.relu(self.bn4(self.conv4(x)))
        x = self.bn5(self.conv5(x))
        if not self.max_pool:
            return x
        x = self.mp1(x).view(-1, 1024)
        if self.global_feat:
            return x, trans
        return torch.cat([x.view(-1, 1024, 1).repeat(1, 1, self.num_points), pointfeat], 1), trans


class PointNetVlad(nn.Module):
    """Reference: PointNetVlad.py:235-247.  forward(x (B,1,N,3)) -> (B, output_dim)."""

    def __init__(self, num_points=2500, global_feat=True, feature_transform=False, max_pool=True, output_dim=1024):
        super().__init__()
        self.point_net = PointNetfeat(num_points=num_points, global_feat=global_feat, feature_transform=feature_transform,
                                      max_pool=max_pool)
        self.net_vlad = NetVLADLoupe(feature_size=1024, max_samples=num_points, cluster_size=64, output_dim=output_dim,
                                     gating=True, add_batch_norm=True, is_training=True)

    def forward(self, x):
        return self.net_vlad(self.point_net(x))
