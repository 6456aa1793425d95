"""PatchAugNet — host-side mirror of ``place_recognition/patch_aug_net/models/patch_aug_net.py``.

``Network(param, use_a2a_recon, use_l2_norm).forward(x, nn_dict=None, return_feat=True)`` keeps the reference's
signature, return structure and ``state_dict`` layout (``backbone.SA_modules.{i}.mlps.0.layer{j}.conv.weight`` ...,
``backbone.FP_modules.{i}.mlp...``, ``aggregation.vlads.{i}...``, ``aggregation.afa...``, ``decoder.fc{1,2,3}``), so
``SceneDataSet.make_descs`` (datasets/scene_dataset.py:675-692) and reference checkpoints work unchanged.

Two execution paths, both on hand-written sm_100a kernels:
  * eval mode on CUDA with the shipped configuration (kNN grouping, ``AGGREGATION_TYPE`` 2, no gating):
    ``patchaugnet_b200.engine.FusedPatchAugNet`` — the whole descriptor extraction in ~25 fused launches;
  * otherwise (training, ``nn_dict`` patch tasks, other aggregation types): the reference's op-by-op structure on
    ``patchaugnet_b200.pointops`` autograd Functions + PyTorch layers.
"""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import loupe as lp
from . import pointops
from . import pt_util
from .pointnet_autoencoder import PointNetDecoder

__all__ = ["Network", "PointNet2", "PointNet2SAModule", "PointNet2SAModuleMSG", "PointNet2FPModule"]


class Network(nn.Module):
    """Reference: patch_aug_net.py:22-107."""

    def __init__(self, param=None, use_a2a_recon=False, use_l2_norm=False):
        super().__init__()
        self.backbone = PointNet2(param=param)
        aggregation = param["AGGREGATION"]
        if aggregation == "spvlad":
            self.aggregation = lp.SpatialPyramidNetVLAD(
                feature_size=param["FEATURE_SIZE"], max_samples=param["MAX_SAMPLES"], cluster_size=param["CLUSTER_SIZE"],
                output_dim=param["OUTPUT_DIM"], gating=param["GATING"], aggregation_type=param["AGGREGATION_TYPE"],
                add_batch_norm=True)
        else:
            print("No aggregation algorithm: ", aggregation)
        self.use_l2_norm = use_l2_norm
        self.use_a2a_recon = use_a2a_recon
        if self.use_a2a_recon:
            self.decoder = PointNetDecoder(embedding_size=256, num_points=param["KNN"][0])
        self._engine = None
        self._engine_key = None
        self.use_fused = True

    # ---- fused eval path ---------------------------------------------------------------------------------------
    def fusable(self):
        agg = getattr(self, "aggregation", None)
        if agg is None or agg.aggregation_type not in (0, 1, 2, 3, 4, 5):
            return False
        for mod in self.backbone.SA_modules:
            if len(mod.groupers) != 1 or mod.npoint is None:
                return False
            g = mod.groupers[0]
            if not isinstance(g, pointops.QueryAndGroup_Edge) or not g.use_xyz:
                return False
            if g.nsample < 2:      # QueryAndGroup_Edge skips the centre subtraction for single-neighbour groups
                return False       # (pointops.py:562-563); the fused loaders always subtract
        return len(self.backbone.SA_modules) == 3 and len(self.backbone.FP_modules) == 3

    def _weights_key(self):
        """Changes whenever any parameter / buffer is modified in place, replaced or moved (optimizer.step(), a parent's
        or a child's load_state_dict, .to(device), frozen-BN fine-tuning in eval mode): tensor identity, device and
        the autograd version counter of every entry."""
        return tuple((id(t), t.device, t._version) for t in list(self.parameters()) + list(self.buffers()))

    def engine(self, refresh=False):
        """The fused inference engine bound to this module's parameters.  Built lazily; the folded weights (BatchNorm
        folded in, bf16 hi/lo planes) are refolded whenever the module's weights changed since they were derived."""
        from .engine import FusedPatchAugNet
        key = self._weights_key()
        if self._engine is None or self._engine.device != key[0][1]:
            self._engine = FusedPatchAugNet(self)
        elif refresh or key != self._engine_key:
            self._engine.refold()
        self._engine_key = key
        return self._engine

    def train(self, mode=True):
        if mode and self._engine is not None:
            self._engine = None        # weights are about to change: drop the folded copy and its workspaces
        return super().train(mode)

    def forward(self, x, nn_dict=None, return_feat=True):
        """x: B x 1 x N x 3"""
        # nn.DataParallel replicas (train_place_recognition.py:546-548) are shallow copies rebuilt on every forward whose
        # parameters() is empty: they take the op-by-op path on the same kernels instead of refolding every call
        if (self.use_fused and not self.training and nn_dict is None and x.is_cuda and not torch.is_grad_enabled()
                and not getattr(self, "_is_replica", False) and self.fusable()):
            with torch.cuda.device(x.device):
                return self.engine()(x, return_feat=return_feat)
        x = x.squeeze(1)
        xyz = x
        res = self.backbone(x)
        center_idx = res["center_idx_origin"]
        sample_idx = res["sample_idx_origin"]
        fp_features = res["fp_features"]
        out = self.aggregation(fp_features)
        if nn_dict is not None:
            # patch reconstruction / augmentation inputs for the first-level patches of the related clouds
            # (reference patch_aug_net.py:68-104)
            related = set()
            for i, j in nn_dict:
                related.add(i)
                related.add(j)
            related = list(related)
            centers, origin_out, feats, recon_out = [], [], [], []
            origin_patches = pointops.grouping(xyz.transpose(1, 2).contiguous(), sample_idx[0])   # B x 3 x M x K
            # The reference selects every related cloud with its own one-element index tensor (host -> device copy, index_select,
            # and in backward a zero-filled full-batch gradient per cloud: patch_aug_net.py:83-98).  One index_select of all
            # related clouds + unbind gives the same per-cloud tensors with one gather forward and one stack backward.
            key = (tuple(related), out.device)
            if getattr(self, "_rel_key", None) != key:                          # cached: no host-to-device copy per step
                self._rel_key, self._rel_idx = key, torch.as_tensor(related, dtype=torch.long, device=out.device)
            rel = self._rel_idx
            f_all = torch.index_select(fp_features[1], 0, rel)                  # R x 256 x M x 1
            p_all = torch.index_select(origin_patches, 0, rel)                  # R x 3 x M x K
            c_all = torch.index_select(center_idx[0], 0, rel)                   # R x M
            for f_r, p_r, c_r in zip(f_all.unbind(0), p_all.unbind(0), c_all.unbind(0)):
                f = f_r.squeeze().transpose(1, 0)                               # M x 256
                if self.use_l2_norm:
                    f = F.normalize(f)
                patches = p_r.transpose(2, 0).transpose(1, 0)                   # M x K x 3
                centers.append(c_r.unsqueeze(0))
                origin_out.append(patches)
                feats.append(f)
                if self.use_a2a_recon:
                    recon_out.append(self.decoder(f))
            out = out, {"cloud_indices": related, "center_indices": centers, "origin_patches": origin_out,
                        "patch_features": feats, "reconstructed_patches": recon_out}
        if return_feat:
            out = out, fp_features, center_idx
        return out


class PointNet2(nn.Module):
    """EdgeConv PointNet++ backbone: 3 set-abstraction + 3 feature-propagation modules.  Reference: patch_aug_net.py:110-192."""

    def __init__(self, param=None):
        super().__init__()
        c = 3
        sap, knn, dil, gp = param["SAMPLING"], param["KNN"], param["KNN_DILATION"], param["GROUP"]
        self.use_origin_pc_in_fp = param["USE_ORIGIN_PC_IN_FP"]
        self.SA_modules = nn.ModuleList([
            PointNet2SAModule(npoint=sap[0], nsample=knn[0], knn_dilation=dil, gp=gp, mlp=[c, 32, 32, 64], use_xyz=True),
            PointNet2SAModule(npoint=sap[1], nsample=knn[1], knn_dilation=dil, gp=gp, mlp=[64, 64, 64, 256], use_xyz=True),
            PointNet2SAModule(npoint=sap[2], nsample=knn[2], knn_dilation=dil, gp=gp, mlp=[256, 256, 256, 512], use_xyz=True),
        ])
        fs = param["FEATURE_SIZE"]
        if not self.use_origin_pc_in_fp:
            c = 0
        self.FP_modules = nn.ModuleList([
            PointNet2FPModule(mlp=[fs[1] + c, 256, 256, fs[0]]),
            PointNet2FPModule(mlp=[fs[2] + 64, 256, fs[1]]),
            PointNet2FPModule(mlp=[512 + 256, 256, fs[2]]),
        ])

    def forward(self, pointcloud):
        l_xyz = [pointcloud]
        l_features = [pointcloud.transpose(1, 2).contiguous()]
        l_center_idx, l_sample_idx = [], []
        for i, sa in enumerate(self.SA_modules):
            xyz_i, cidx_i, sidx_i, feat_i = sa(l_xyz[i], l_features[i])
            l_xyz.append(xyz_i)
            l_features.append(feat_i)
            l_center_idx.append(cidx_i)
            l_sample_idx.append(sidx_i)
        sa_features = list(l_features)
        # indices expressed in the ORIGINAL cloud (patch_aug_net.py:169-177)
        c_origin, s_origin = [l_center_idx[0]], [l_sample_idx[0]]
        for i in range(1, len(l_center_idx)):
            c_origin.append(torch.gather(c_origin[i - 1], -1, l_center_idx[i].long()))
            table = c_origin[i - 1].unsqueeze(1).repeat(1, l_sample_idx[i].shape[1], 1)
            s_origin.append(torch.gather(table, -1, l_sample_idx[i].long()))
        nfp = len(self.FP_modules)
        for i in range(-1, -(nfp + 1), -1):
            skip = None if (i == -nfp and not self.use_origin_pc_in_fp) else l_features[i - 1]
            l_features[i - 1] = self.FP_modules[i](l_xyz[i - 1], l_xyz[i], skip, l_features[i])
        return {"center_idx_origin": c_origin, "sample_idx_origin": s_origin,
                "sa_features": [sa_features[1], sa_features[2], sa_features[3]],
                "fp_features": [l_features[2].unsqueeze(-1), l_features[1].unsqueeze(-1), l_features[0].unsqueeze(-1)]}


class _PointNet2SAModuleBase(nn.Module):
    """FPS -> gather centres -> group -> SharedMLP -> max over K.  Reference: patch_aug_net.py:195-243."""

    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None

    def forward(self, xyz, features=None):
        xyz_trans = xyz.transpose(1, 2).contiguous()
        center_idx = pointops.furthestsampling(xyz, self.npoint)
        new_xyz = pointops.gathering(xyz_trans, center_idx).transpose(1, 2).contiguous() if self.npoint is not None else None
        center_features = pointops.gathering(features, center_idx)
        outs, sidx = [], []
        for grouper, mlp in zip(self.groupers, self.mlps):
            new_features, sample_idx = grouper(xyz, new_xyz, features, center_features)   # B x C x M x K
            new_features = mlp(new_features)
            new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)]).squeeze(-1)
            outs.append(new_features)
            sidx.append(sample_idx)
        return new_xyz, center_idx, torch.cat(sidx, dim=-1), torch.cat(outs, dim=1)


class PointNet2SAModuleMSG(_PointNet2SAModuleBase):
    """Multi-scale grouping variant.  Reference: patch_aug_net.py:246-290."""

    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int], knn_dilation: int, mlps: List[List[int]],
                 gp: int, bn: bool = True, use_xyz: bool = True):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for radius, nsample, spec in zip(radii, nsamples, mlps):
            self.groupers.append(
                pointops.QueryAndGroup_Edge(radius, nsample, knn_dilation=knn_dilation, use_xyz=use_xyz, ret_sample_idx=True)
                if npoint is not None else pointops.GroupAll(use_xyz))
            if use_xyz:
                spec[0] += 3
            self.mlps.append(pt_util.SharedMLP(spec, bn=bn))


class PointNet2SAModule(PointNet2SAModuleMSG):
    """Single-scale set abstraction.  Reference: patch_aug_net.py:293-314."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 knn_dilation: int = 1, gp: int = None, bn: bool = True, use_xyz: bool = True):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], knn_dilation=knn_dilation, gp=gp,
                         bn=bn, use_xyz=use_xyz)


class PointNet2FPModule(nn.Module):
    """3-NN inverse-distance interpolation + skip concat + SharedMLP.  Reference: patch_aug_net.py:317-363."""

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
