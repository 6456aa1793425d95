"""Torch-CPU restatement of the reference's PatchAugNet eval forward — TEST INFRASTRUCTURE (see oracle/__init__.py).

A functional forward over a reference-layout ``state_dict``: the point-cloud primitives go through the C oracle
(``oracle.ops``), the dense layers through plain ``torch`` CPU ops in the reference's own sequence
(conv -> BatchNorm(eval) -> ReLU as separate steps, no folding).  ``dtype=torch.float64`` runs the dense part in
double (indices are still decided in fp32 exactly like the reference) to serve as a tighter ground truth.

Follows (paths relative to /root/reference):
  place_recognition/patch_aug_net/models/patch_aug_net.py:48-107, 141-192, 203-243, 331-363
  libs/pointops/functions/pointops.py:533-582
  utils/model_util/pt_util.py:98-151
  place_recognition/patch_aug_net/models/loupe.py:24-41, 57-66, 191-222, 278-329

Pinned by tests/golden/patchaugnet_ref_forward.npz, produced by running the reference's own nn.Module code in the
build container (tests/golden/make_golden.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import ops

BN_EPS = 1e-5


def _bn(x, sd, prefix, dim):
    """eval-mode BatchNorm on channel dim `dim` (running statistics)."""
    shape = [1] * x.dim()
    shape[dim] = -1
    g, b = sd[prefix + ".weight"].to(x.dtype), sd[prefix + ".bias"].to(x.dtype)
    mu, var = sd[prefix + ".running_mean"].to(x.dtype), sd[prefix + ".running_var"].to(x.dtype)
    return (x - mu.view(shape)) / torch.sqrt(var.view(shape) + BN_EPS) * g.view(shape) + b.view(shape)


def shared_mlp(x, sd, prefix):
    """pt_util.SharedMLP on (B,C,M,K): per layer conv1x1 (no bias) -> BN2d -> ReLU.  pt_util.py:16-41, 98-151."""
    i = 0
    while f"{prefix}.layer{i}.conv.weight" in sd:
        w = sd[f"{prefix}.layer{i}.conv.weight"].to(x.dtype)
        x = F.conv2d(x, w)
        x = _bn(x, sd, f"{prefix}.layer{i}.bn.bn", 1)
        x = F.relu(x)
        i += 1
    return x


def _t(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def sa_module(xyz, features, sd, prefix, npoint, nsample, dilation, perm, dtype):
    """_PointNet2SAModuleBase.forward (patch_aug_net.py:203-243) with QueryAndGroup_Edge (pointops.py:533-582).
    xyz (B,n,3) float32 numpy; features (B,C,n) torch.  Returns new_xyz, center_idx, sample_idx, new_features."""
    center_idx = ops.furthestsampling(xyz, npoint)                                   # :220
    new_xyz = np.take_along_axis(xyz, center_idx[..., None].astype(np.int64), 1)    # :222-225
    center_features = torch.gather(features, 2, _t(center_idx, torch.int64)[:, None, :].expand(-1, features.shape[1], -1))
    if dilation > 1:                                                                  # pointops.py:551-555
        cand = ops.knnquery(dilation * nsample, xyz, new_xyz)
        if perm is None:
            perm = torch.randperm(nsample).numpy()
        idx = np.ascontiguousarray(cand[:, :, perm])
    else:
        idx = ops.knnquery(nsample, xyz, new_xyz)
    B, m, k = idx.shape
    li = _t(idx, torch.int64).view(B, 1, m * k)
    xyz_t = _t(xyz).transpose(1, 2)                                                   # B x 3 x n (fp32, like the reference)
    grouped_xyz = torch.gather(xyz_t, 2, li.expand(-1, 3, -1)).view(B, 3, m, k) - _t(new_xyz).transpose(1, 2).unsqueeze(-1)
    grouped_feat = torch.gather(features, 2, li.expand(-1, features.shape[1], -1)).view(B, -1, m, k) - center_features.unsqueeze(-1)
    new_features = torch.cat([grouped_xyz.to(dtype), grouped_feat.to(dtype)], dim=1)  # :570
    new_features = shared_mlp(new_features, sd, prefix + ".mlps.0")                   # patch_aug_net.py:235
    new_features = new_features.max(dim=3)[0]                                         # :236-237
    return new_xyz, center_idx, idx, new_features


def fp_module(unknown, known, unknown_feats, known_feats, sd, prefix, dtype):
    """PointNet2FPModule.forward (patch_aug_net.py:331-363).  Weights are computed in fp32 like the reference."""
    d2, idx = ops.nearestneighbor(unknown, known)
    dist = torch.sqrt(_t(d2))                                                         # pointops.py:76
    dist_recip = 1.0 / (dist + 1e-8)
    weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)                  # :351-353
    if known_feats.dtype == torch.float32:
        interp = _t(ops.interpolation(known_feats.numpy(), idx, weight.numpy()))      # exact fma order of the kernel
    else:
        B, c, m = known_feats.shape
        n = idx.shape[1]
        g = torch.gather(known_feats, 2, _t(idx, torch.int64).view(B, 1, n * 3).expand(-1, c, -1)).view(B, c, n, 3)
        interp = (g * weight.to(dtype).unsqueeze(1)).sum(-1)
    new = torch.cat([interp, unknown_feats.to(dtype)], dim=1) if unknown_feats is not None else interp
    return shared_mlp(new.unsqueeze(-1), sd, prefix + ".mlp").squeeze(-1)


def netvlad(x, sd, prefix):
    """NetVLADBase.forward, loupe.py:191-222.  x (B,C,N,1) -> (B,C,K)."""
    B, C, N, _ = x.shape
    cw = sd[prefix + ".cluster_weights"].to(x.dtype)
    cw2 = sd[prefix + ".cluster_weights2"].to(x.dtype)
    K = cw.shape[1]
    x = x.transpose(1, 3).contiguous().view(-1, N, C)
    act = torch.matmul(x, cw)
    act = _bn(act.view(-1, K), sd, prefix + ".bn1", 1).view(-1, N, K)
    act = torch.softmax(act, dim=-1)
    a = act.sum(-2, keepdim=True) * cw2
    vlad = torch.matmul(act.transpose(2, 1), x).transpose(2, 1) - a
    return F.normalize(vlad, dim=1, p=2).contiguous()


def afa(x, sd, prefix):
    """AdaptiveFeatureAggregator.forward (loupe.py:57-66) with MLPAttentionLayer 'way 2' (loupe.py:24-41)."""
    res = x
    i = 0
    while f"{prefix}.mlpa.mlps.{i}.weight" in sd:
        res = F.conv1d(res, sd[f"{prefix}.mlpa.mlps.{i}.weight"].to(x.dtype))
        i += 1
    w = torch.softmax(res.max(dim=1)[0], dim=-1).unsqueeze(1)
    x = F.relu(x + x * w)
    B, C, K = x.shape
    x = F.linear(x.reshape(B, C * K), sd[prefix + ".fc.weight"].to(x.dtype), sd[prefix + ".fc.bias"].to(x.dtype))
    x = _bn(x, sd, prefix + ".bn", 1)
    return F.normalize(x)


def patchaugnet_forward(sd, cfg, x, perms=None, dtype=torch.float32):
    """Network.forward(x) with nn_dict=None, return_feat=True (patch_aug_net.py:48-107), eval mode.

    x: (B,1,N,3) or (B,N,3) float32 array/tensor.  perms: optional list of 3 permutations (the torch.randperm draws).
    Returns dict(desc (B,256), fp_features [3], center_idx_origin [3], sample_idx_origin [3], sa_features [3], xyz [4]).
    """
    xyz0 = np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(np.asarray(x).shape[0], -1, 3))
    sd = {k: (v.detach().cpu() if torch.is_tensor(v) else torch.as_tensor(v)) for k, v in sd.items()}
    sap, knn, dil = cfg["SAMPLING"], cfg["KNN"], cfg["KNN_DILATION"]
    l_xyz = [xyz0]
    l_feat = [_t(xyz0).transpose(1, 2).contiguous().to(dtype)]
    cidx, sidx = [], []
    for i in range(3):
        nx, ci, si, nf = sa_module(l_xyz[i], l_feat[i], sd, f"backbone.SA_modules.{i}", sap[i], knn[i], dil,
                                   None if perms is None else perms[i], dtype)
        l_xyz.append(nx); l_feat.append(nf); cidx.append(ci); sidx.append(si)
    sa_features = list(l_feat[1:])
    c_origin, s_origin = [cidx[0]], [sidx[0]]
    for i in range(1, 3):                                                              # patch_aug_net.py:169-177
        c_origin.append(np.take_along_axis(c_origin[i - 1], cidx[i].astype(np.int64), -1))
        table = np.repeat(c_origin[i - 1][:, None, :], sidx[i].shape[1], 1)
        s_origin.append(np.take_along_axis(table, sidx[i].astype(np.int64), -1))
    use_origin = cfg["USE_ORIGIN_PC_IN_FP"]
    for i in range(-1, -4, -1):                                                        # :183-187
        skip = None if (i == -3 and not use_origin) else l_feat[i - 1]
        l_feat[i - 1] = fp_module(l_xyz[i - 1], l_xyz[i], skip, l_feat[i], sd, f"backbone.FP_modules.{3 + i}", dtype)
    fp_features = [l_feat[2].unsqueeze(-1), l_feat[1].unsqueeze(-1), l_feat[0].unsqueeze(-1)]
    v = torch.cat([netvlad(f, sd, f"aggregation.vlads.{i}") for i, f in enumerate(fp_features)], dim=-1)
    assert cfg["AGGREGATION_TYPE"] == 2 and not cfg["GATING"], "oracle restates the shipped configuration only"
    desc = afa(v, sd, "aggregation.afa")
    return dict(desc=desc, fp_features=fp_features, center_idx_origin=c_origin, sample_idx_origin=s_origin,
                sa_features=sa_features, xyz=l_xyz, vlad=v)


# ---- PPT-Net -------------------------------------------------------------------------------------------------------
# place_recognition/pptnet_origin/models/pptnet.py:46-62, 90-134, 145-183, 246-282, 313-330
# place_recognition/pptnet_origin/models/loupe.py:39-71, 94-105, 124-136

def sa_layer(x, sd, prefix, gp):
    """SA_Layer.forward (pptnet.py:261-282), eval mode.  x (B,C,N).  q and k share one weight (pptnet.py:254)."""
    B, C, N = x.shape
    wk = sd[prefix + ".k_conv.weight"].to(x.dtype)
    q = F.conv1d(x, wk, groups=gp).reshape(B, gp, C // gp, N).permute(0, 1, 3, 2)
    k = F.conv1d(x, wk, groups=gp).reshape(B, gp, C // gp, N)
    v = F.conv1d(x, sd[prefix + ".v_conv.weight"].to(x.dtype), sd[prefix + ".v_conv.bias"].to(x.dtype))
    energy = torch.matmul(q, k).sum(dim=1)
    attn = torch.softmax(energy, dim=-1)
    attn = attn / (1e-9 + attn.sum(dim=1, keepdim=True))
    x_r = torch.matmul(v, attn)
    t = F.conv1d(x - x_r, sd[prefix + ".trans_conv.weight"].to(x.dtype), sd[prefix + ".trans_conv.bias"].to(x.dtype))
    return x + F.relu(_bn(t, sd, prefix + ".after_norm", 1))


def gating_context(x, sd, prefix):
    """GatingContext.forward, pptnet_origin/models/loupe.py:124-136 (add_batch_norm=True)."""
    gates = torch.matmul(x, sd[prefix + ".gating_weights"].to(x.dtype))
    gates = _bn(gates, sd, prefix + ".bn1", 1)
    return x * torch.sigmoid(gates)


def pptnet_forward(sd, cfg, x, dtype=torch.float32, use_normalize=True):
    """pptnet.Network.forward(x), eval mode.  Returns dict(desc, fp_features [4], center_idx_origin [4])."""
    xyz0 = np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(np.asarray(x).shape[0], -1, 3))
    sd = {k: (v.detach().cpu() if torch.is_tensor(v) else torch.as_tensor(v)) for k, v in sd.items()}
    sap, knn, gp = cfg["SAMPLING"], cfg["KNN"], cfg["GROUP"]
    l_xyz = [xyz0]
    l_feat = [_t(xyz0).transpose(1, 2).contiguous().to(dtype)]
    cidx = []
    for i in range(4):
        nx, ci, si, nf = sa_module(l_xyz[i], l_feat[i], sd, f"backbone.SA_modules.{i}", sap[i], knn[i], 1, None, dtype)
        nf = sa_layer(nf, sd, f"backbone.SA_modules.{i}.sas.0", gp)
        l_xyz.append(nx); l_feat.append(nf); cidx.append(ci)
    c_origin = [cidx[0]]
    for i in range(1, 4):
        c_origin.append(np.take_along_axis(c_origin[i - 1], cidx[i].astype(np.int64), -1))
    for i in range(-1, -5, -1):
        l_feat[i - 1] = fp_module(l_xyz[i - 1], l_xyz[i], l_feat[i - 1], l_feat[i], sd, f"backbone.FP_modules.{4 + i}", dtype)
    fp_features = [l_feat[3].unsqueeze(-1), l_feat[2].unsqueeze(-1), l_feat[1].unsqueeze(-1), l_feat[0].unsqueeze(-1)]
    vs = []
    for i, f in enumerate(fp_features):
        v = netvlad(f, sd, f"aggregation.vlad{i}")                       # (B,C,K)
        vs.append(v.reshape(v.shape[0], -1))                             # view(B, C*K), loupe.py:70
    vlad = torch.cat(vs, dim=-1)
    vlad = torch.matmul(vlad, sd["aggregation.hidden_weights"].to(dtype))
    vlad = _bn(vlad, sd, "aggregation.bn2", 1)
    if cfg["GATING"]:
        vlad = gating_context(vlad, sd, "aggregation.context_gating")
    if use_normalize:
        vlad = F.normalize(vlad)
    return dict(desc=vlad, fp_features=fp_features, center_idx_origin=c_origin, xyz=l_xyz)
