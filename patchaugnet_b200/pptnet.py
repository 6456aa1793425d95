"""PPT-Net — host-side mirror of ``place_recognition/pptnet_origin/models/pptnet.py``.

``Network(param, use_normalize).forward(x, return_feat=True)`` keeps the reference signature, return structure and
``state_dict`` layout (258 entries; ``sas.0.q_conv.weight`` and ``sas.0.k_conv.weight`` are the SAME tensor, pptnet.py:254).
Backbone: 4 set-abstraction modules, each followed by the grouped self-attention ``SA_Layer`` (pptnet.py:246-282),
4 feature-propagation modules, 4-level NetVLAD + fc + gating head.  All point-cloud primitives run on the
hand-written kernels (``patchaugnet_b200.pointops``); ``SA_Layer`` in eval mode on CUDA runs the fused attention
kernels (``patchaugnet_b200.attention``), otherwise the reference's sequence of torch ops.
"""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointops
from . import pptnet_loupe as lp
from . import pt_util

__all__ = ["Network", "PointNet2", "SA_Layer", "PointNet2SAModule", "PointNet2FPModule"]


class Network(nn.Module):
    """Reference: pptnet.py:24-62."""

    def __init__(self, param=None, use_normalize=True):
        super().__init__()
        self.backbone = PointNet2(param=param)
        aggregation = param["AGGREGATION"]
        if aggregation == "spvlad":
            self.aggregation = lp.SpatialPyramidNetVLAD(
                feature_size=param["FEATURE_SIZE"], max_samples=param["MAX_SAMPLES"], cluster_size=param["CLUSTER_SIZE"],
                output_dim=param["OUTPUT_DIM"], gating=param["GATING"], add_batch_norm=True)
        else:
            print("No aggregation algorithm: ", aggregation)
        self.use_normalize = use_normalize
        self.use_fused = True          # eval-mode CUDA forward on the fused engine (engine_ppt.FusedPPTNet)
        self.compute_dtype = "f32"     # "bf16": single-rounding bf16 tensor-core operands in the fused engine (configs[2])
        self._engine = None
        self._engine_key = None

    def engine(self, refresh=False):
        """The fused eval engine; rebuilt when parameters / buffers changed (state_dict load, in-place updates)."""
        key = (self.compute_dtype,) + tuple(p._version for p in self.parameters()) + tuple(b._version for b in self.buffers())
        if self._engine is None or refresh or key != self._engine_key:
            from .engine_ppt import FusedPPTNet
            self._engine = FusedPPTNet(self, precision=self.compute_dtype)
            self._engine_key = key
        return self._engine

    def train(self, mode=True):
        self._engine = None
        return super().train(mode)

    def fusable(self):
        """True when the eval forward runs on the fused engine (what retrieval.extract_descriptors checks before it uses the
        engine's throughput mode)."""
        return bool(self.use_fused) and isinstance(self.aggregation, lp.SpatialPyramidNetVLAD)

    def forward(self, x, return_feat=True):
        """x: B x 1 x N x 3"""
        if (self.use_fused and not self.training and x.is_cuda and not torch.is_grad_enabled()
                and isinstance(self.aggregation, lp.SpatialPyramidNetVLAD)):
            return self.engine()(x, return_feat=return_feat)
        x = x.squeeze(1)
        res = self.backbone(x)
        f = res["fp_features"]
        out = self.aggregation(f[0], f[1], f[2], f[3])
        if self.use_normalize:
            out = F.normalize(out)
        return (out, res["fp_features"], res["center_idx_origin"]) if return_feat else out


class PointNet2(nn.Module):
    """Reference: pptnet.py:65-134."""

    def __init__(self, param=None):
        super().__init__()
        c = 3
        sap, knn, fs, gp = param["SAMPLING"], param["KNN"], param["FEATURE_SIZE"], param["GROUP"]
        self.SA_modules = nn.ModuleList([
            PointNet2SAModule(npoint=sap[0], nsample=knn[0], gp=gp, mlp=[c, 32, 32, 64], use_xyz=True),
            PointNet2SAModule(npoint=sap[1], nsample=knn[1], gp=gp, mlp=[64, 64, 64, 128], use_xyz=True),
            PointNet2SAModule(npoint=sap[2], nsample=knn[2], gp=gp, mlp=[128, 128, 128, 256], use_xyz=True),
            PointNet2SAModule(npoint=sap[3], nsample=knn[3], gp=gp, mlp=[256, 256, 256, 512], use_xyz=True),
        ])
        self.FP_modules = nn.ModuleList([
            PointNet2FPModule(mlp=[fs[1] + c, 256, 256, fs[0]]),
            PointNet2FPModule(mlp=[fs[2] + 64, 256, fs[1]]),
            PointNet2FPModule(mlp=[fs[3] + 128, 256, fs[2]]),
            PointNet2FPModule(mlp=[512 + 256, 256, fs[3]]),
        ])

    def forward(self, pointcloud):
        l_xyz, l_features = [pointcloud], [pointcloud.transpose(1, 2).contiguous()]
        l_center_idx, l_sample_idx = [], []
        for i, sa in enumerate(self.SA_modules):
            xyz_i, cidx_i, sidx_i, feat_i = sa(l_xyz[i], l_features[i])
            l_xyz.append(xyz_i)
            l_features.append(feat_i)
            l_center_idx.append(cidx_i)
            l_sample_idx.append(sidx_i)
        c_origin, s_origin = [l_center_idx[0]], [l_sample_idx[0]]
        for i in range(1, len(l_center_idx)):                                  # pptnet.py:109-118
            c_origin.append(torch.gather(c_origin[i - 1], -1, l_center_idx[i].long()))
            table = c_origin[i - 1].unsqueeze(1).repeat(1, l_sample_idx[i].shape[1], 1)
            s_origin.append(torch.gather(table, -1, l_sample_idx[i].long()))
        for i in range(-1, -(len(self.FP_modules) + 1), -1):
            l_features[i - 1] = self.FP_modules[i](l_xyz[i - 1], l_xyz[i], l_features[i - 1], l_features[i])
        return {"center_idx_origin": c_origin, "sample_idx_origin": s_origin,
                "fp_features": [l_features[3].unsqueeze(-1), l_features[2].unsqueeze(-1), l_features[1].unsqueeze(-1),
                                l_features[0].unsqueeze(-1)]}


class _PointNet2SAModuleBase(nn.Module):
    """FPS -> group -> SharedMLP -> max over K -> SA_Layer.  Reference: pptnet.py:137-183."""

    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None
        self.sas = None

    def forward(self, xyz, features=None):
        xyz_trans = xyz.transpose(1, 2).contiguous()
        center_idx = pointops.furthestsampling(xyz, self.npoint)
        new_xyz = pointops.gathering(xyz_trans, center_idx).transpose(1, 2).contiguous() if self.npoint is not None else None
        center_features = pointops.gathering(features, center_idx)
        outs, sidx = [], []
        for grouper, mlp, sa in zip(self.groupers, self.mlps, self.sas):
            new_features, sample_idx = grouper(xyz, new_xyz, features, center_features)
            new_features = mlp(new_features)
            new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)]).squeeze(-1)
            outs.append(sa(new_features))
            sidx.append(sample_idx)
        return new_xyz, center_idx, torch.cat(sidx, dim=-1), torch.cat(outs, dim=1)


class PointNet2SAModuleMSG(_PointNet2SAModuleBase):
    """Reference: pptnet.py:186-224."""

    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int], mlps: List[List[int]], gp: int,
                 bn: bool = True, use_xyz: bool = True):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        self.sas = nn.ModuleList()
        for radius, nsample, spec in zip(radii, nsamples, mlps):
            self.groupers.append(pointops.QueryAndGroup_Edge(radius, nsample, use_xyz=use_xyz, ret_sample_idx=True)
                                 if npoint is not None else pointops.GroupAll(use_xyz))
            if use_xyz:
                spec[0] += 3
            self.mlps.append(pt_util.SharedMLP(spec, bn=bn))
            self.sas.append(SA_Layer(spec[-1], gp))


class PointNet2SAModule(PointNet2SAModuleMSG):
    """Reference: pptnet.py:227-243."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None, gp: int = None,
                 bn: bool = True, use_xyz: bool = True):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], gp=gp, bn=bn, use_xyz=use_xyz)


class SA_Layer(nn.Module):
    """Grouped self-attention with tied q/k projection.  Reference: pptnet.py:246-282.

    energy = sum over the gp groups of q_g^T k_g = Q^T Q (a Gram matrix, because q and k share weights and the per-group
    dot products add up to the dot product over all channels); attn = row-softmax(energy) / (1e-9 + column sums);
    x_r = V attn;  out = x + relu(BN(trans_conv(x - x_r))).
    """

    def __init__(self, channels, gp):
        super().__init__()
        self.gp = gp
        assert channels % 4 == 0
        self.q_conv = nn.Conv1d(channels, channels, 1, bias=False, groups=gp)
        self.k_conv = nn.Conv1d(channels, channels, 1, bias=False, groups=gp)
        self.q_conv.weight = self.k_conv.weight
        self.v_conv = nn.Conv1d(channels, channels, 1)
        self.trans_conv = nn.Conv1d(channels, channels, 1)
        self.after_norm = nn.BatchNorm1d(channels)
        self.act = nn.ReLU()
        self.softmax = nn.Softmax(dim=-1)
        self.use_fused = True

    def forward(self, x):
        """x: B x C x N"""
        if self.use_fused and not self.training and x.is_cuda and not torch.is_grad_enabled():
            from . import attention
            return attention.sa_layer_forward(self, x, getattr(self, "attention_precision", 2))
        bs, ch, nums = x.size()
        x_q = self.q_conv(x).reshape(bs, self.gp, ch // self.gp, nums).permute(0, 1, 3, 2)
        x_k = self.k_conv(x).reshape(bs, self.gp, ch // self.gp, nums)
        x_v = self.v_conv(x)
        energy = torch.sum(torch.matmul(x_q, x_k), dim=1, keepdims=False)
        attn = self.softmax(energy)
        attn = attn / (1e-9 + attn.sum(dim=1, keepdims=True))
        x_r = torch.matmul(x_v, attn)
        x_r = self.act(self.after_norm(self.trans_conv(x - x_r)))
        return x + x_r


class PointNet2FPModule(nn.Module):
    """Reference: pptnet.py:285-330."""

    def __init__(self, *, mlp: List[int], bn: bool = True):
        super().__init__()
        self.mlp = pt_util.SharedMLP(mlp, bn=bn)

    def forward(self, unknown, known, unknow_feats, known_feats):
        if known is not None:
            dist, idx = pointops.nearestneighbor(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
            interpolated = pointops.interpolation(known_feats, idx, weight)
        else:
            interpolated = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        new_features = torch.cat([interpolated, unknow_feats], dim=1) if unknow_feats is not None else interpolated
        return self.mlp(new_features.unsqueeze(-1)).squeeze(-1)
