"""PointNetDecoder — mirror of ``place_recognition/patch_aug_net/models/pointnet_autoencoder.py:85-111``.

fc 256 -> 1024 -> 1024 -> num_points*3 with BN1d + ReLU, tanh output, used by the patch-reconstruction task
(patch_aug_net.py:46, 97).  Same parameter names (fc1, fc2, bn1, bn2, fc3) as the reference.
"""
import torch
import torch.nn.functional as F
from torch import nn


class PointNetDecoder(nn.Module):
    def __init__(self, embedding_size, output_channels=3, num_points=1024):
        super().__init__()
        self.num_points = num_points
        self.output_channels = output_channels
        self.fc1 = nn.Linear(embedding_size, 1024)
        self.fc2 = nn.Linear(1024, 1024)
        self.bn1 = nn.BatchNorm1d(1024)
        self.bn2 = nn.BatchNorm1d(1024)
        self.fc3 = nn.Linear(1024, num_points * output_channels)

    def forward(self, x):
        """x: (B, C) -> (B, num_points, 3)"""
        x = F.relu(self.bn1(self.fc1(x)))
        x = F.relu(self.bn2(self.fc2(x)))
        x = torch.tanh(self.fc3(x))
        return x.view(x.shape[0], self.num_points, self.output_channels).contiguous()
