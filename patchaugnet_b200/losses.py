"""Losses used by the training step — mirror of ``losses/pointnetvlad_loss.py`` (the functions the shipped configs reach:
``quadruplet_loss`` :53-105, ``triplet_loss`` :18-45, ``contrastive_loss`` :170-186, ``chamfer_loss`` :189-203,
``patch_chamfer_loss`` :242-247, ``emd_loss`` :205-221, ``patch_emd_loss`` :250-256).  The descriptor losses are plain
torch (negligible cost); the point-set losses run on the hand-written chamfer / EMD kernels.
"""
import torch
import torch.nn.functional as F

from .chamfer_dist import ChamferDistanceL1
from .emd_module import emdModule


def best_pos_distance(query, pos_vecs):
    """Squared distance from each query (A,1,D) to its closest / farthest positive (A,P,D).  Reference :9-15."""
    diff = ((pos_vecs - query.repeat(1, int(pos_vecs.shape[1]), 1)) ** 2).sum(2)
    return diff.min(1)[0], diff.max(1)[0]


def _reduce(loss, lazy, ignore_zero_loss, sum_not_mean=False):
    loss = loss.max(1)[0] if lazy else (loss.sum(1) if sum_not_mean else loss.mean(1))
    if ignore_zero_loss:
        hard = torch.gt(loss, 1e-16).float().sum()
        return loss.sum() / (hard + 1e-16)
    return loss.mean()


def triplet_loss(q_vec, pos_vecs, neg_vecs, margin, use_min=False, lazy=False, ignore_zero_loss=False):
    min_pos, max_pos = best_pos_distance(q_vec, pos_vecs)
    positive = (min_pos if use_min else max_pos).view(-1, 1).repeat(1, int(neg_vecs.shape[1]))
    loss = (margin + positive - ((neg_vecs - q_vec.repeat(1, int(neg_vecs.shape[1]), 1)) ** 2).sum(2)).clamp(min=0.0)
    return _reduce(loss, lazy, ignore_zero_loss, sum_not_mean=True)


def triplet_loss_wrapper(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2, use_min=False, lazy=False, ignore_zero_loss=False):
    return triplet_loss(q_vec, pos_vecs, neg_vecs, m1, use_min, lazy, ignore_zero_loss)


def quadruplet_loss(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2, use_min=False, lazy=False, ignore_zero_loss=False,
                    soft_margin=False):
    """Lazy quadruplet loss of PointNetVLAD as the reference computes it (max_pos unless use_min)."""
    min_pos, max_pos = best_pos_distance(q_vec, pos_vecs)
    num_neg = int(neg_vecs.shape[1])
    positive = (min_pos if use_min else max_pos).view(-1, 1).repeat(1, num_neg)

    def hinge(x):
        return torch.log(1 + torch.exp(x.clamp(max=88))) if soft_margin else x.clamp(min=0.0)

    first = hinge(m1 + positive - ((neg_vecs - q_vec.repeat(1, num_neg, 1)) ** 2).sum(2))
    second = hinge(m2 + positive - ((neg_vecs - other_neg.repeat(1, num_neg, 1)) ** 2).sum(2))
    return _reduce(first, lazy, ignore_zero_loss) + _reduce(second, lazy, ignore_zero_loss)


def contrastive_loss(q_vec, pos_vec, neg_vec, margin):
    """Lists of (D,) feature tensors.  Reference :170-186."""
    total = 0.0
    q = torch.stack(q_vec, dim=0)
    if len(pos_vec) > 0:
        total = total + torch.mean(torch.pow(F.pairwise_distance(q, torch.stack(pos_vec, dim=0)), 2))
    if len(neg_vec) > 0:
        d = F.pairwise_distance(q, torch.stack(neg_vec, dim=0))
        total = total + torch.mean(torch.pow(torch.clamp(margin - d, min=0.0), 2))
    return total


def chamfer_loss(pc1, pc2):
    a = torch.cat([p.float() for p in pc1], 1).squeeze(0)
    b = torch.cat([p.float() for p in pc2], 1).squeeze(0)
    return ChamferDistanceL1()(a, b)


def patch_chamfer_loss(origin_patches, recon_patches):
    """(n_patches, 20, 3) patch sets of all related clouds.  Reference :242-247."""
    return ChamferDistanceL1()(torch.cat(origin_patches, 0), torch.cat(recon_patches, 0))


def emd_loss(pc1, pc2):
    a = torch.cat([p.float() for p in pc1], 1).view(-1, 4096, 3)
    b = torch.cat([p.float() for p in pc2], 1).view(-1, 4096, 3)
    dis, _ = emdModule()(a, b, 0.02, 1024)
    return torch.mean(torch.mean(torch.sqrt(dis), dim=1))


def patch_emd_loss(origin_patches, recon_patches):
    """Reference :250-256.  NOTE: the reference's EMD kernel rejects n % 1024 != 0 and its wrapper ignores the error
    (SURVEY.md note B), so this path silently returns 0 there; here the unsupported shape raises."""
    dis, _ = emdModule()(torch.cat(origin_patches, 0), torch.cat(recon_patches, 0), 0.02, 1024)
    return torch.mean(torch.mean(torch.sqrt(dis), dim=1))
