"""Host-side mirror of the reference's operator module ``libs/pointops/functions/pointops.py``.

Same public names, argument meaning, return values and autograd behaviour (index-producing ops return ``None``
gradients; gather / group / interpolate scatter their gradients back), so model code written against the reference
module runs unchanged:

    furthestsampling, gathering, nearestneighbor, interpolation, grouping, grouping_int, ballquery,
    featuredistribute, featuregather, labelstat_ballrange, labelstat_idx, labelstat_and_ballquery,
    pairwise_distances, knnquery_naive, knnquery, knnquery_exclude,
    QueryAndGroup, QueryAndGroup_Edge, QueryAndGroup_Edge_Split, GroupAll

Every CUDA op goes through ``patchaugnet_b200.pointops_cuda`` (the C ABI of libpatchaug_b200.so) on the current
stream; outputs are allocated on the INPUT's device (the reference uses ``torch.cuda.FloatTensor(...)``, i.e. the
current device, pointops.py:20-21).  There is no CPU path.
"""
from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from . import pointops_cuda as K


def _new(ref: torch.Tensor, shape, dtype, zero=False):
    return (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=ref.device)


class FurthestSampling(Function):
    """xyz (b,n,3), m -> idx (b,m) int32.  Reference: pointops.py:11-29."""

    @staticmethod
    def forward(ctx, xyz, m):
        assert xyz.is_contiguous()
        b, n, _ = xyz.size()
        idx = _new(xyz, (b, m), torch.int32)
        temp = torch.full((b, n), 1e10, dtype=torch.float32, device=xyz.device)
        K.furthestsampling_cuda(b, n, m, xyz, temp, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthestsampling = FurthestSampling.apply


class Gathering(Function):
    """features (b,c,n), idx (b,m) -> (b,c,m).  Reference: pointops.py:32-57."""

    @staticmethod
    def forward(ctx, features, idx):
        assert features.is_contiguous() and idx.is_contiguous()
        b, c, n = features.size()
        m = idx.size(1)
        out = _new(features, (b, c, m), torch.float32)
        K.gathering_forward_cuda(b, c, n, m, features, idx, out)
        ctx.for_backwards = (idx, c, n)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, c, n = ctx.for_backwards
        b, m = idx.size()
        grad = _new(grad_out, (b, c, n), torch.float32, zero=True)
        K.gathering_backward_cuda(b, c, n, m, grad_out.detach().contiguous(), idx, grad)
        return grad, None


gathering = Gathering.apply


class NearestNeighbor(Function):
    """unknown (b,n,3), known (b,m,3) -> (sqrt dist (b,n,3), idx (b,n,3)).  Reference: pointops.py:60-82."""

    @staticmethod
    def forward(ctx, unknown, known) -> Tuple[torch.Tensor, torch.Tensor]:
        assert unknown.is_contiguous() and known.is_contiguous()
        b, n, _ = unknown.size()
        m = known.size(1)
        dist2 = _new(unknown, (b, n, 3), torch.float32)
        idx = _new(unknown, (b, n, 3), torch.int32)
        K.nearestneighbor_cuda(b, n, m, unknown, known, dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


nearestneighbor = NearestNeighbor.apply


class Interpolation(Function):
    """features (b,c,m), idx (b,n,3), weight (b,n,3) -> (b,c,n).  Reference: pointops.py:85-118."""

    @staticmethod
    def forward(ctx, features, idx, weight):
        assert features.is_contiguous() and idx.is_contiguous() and weight.is_contiguous()
        b, c, m = features.size()
        n = idx.size(1)
        ctx.interpolation_for_backward = (idx, weight, m)
        out = _new(features, (b, c, n), torch.float32)
        K.interpolation_forward_cuda(b, c, m, n, features, idx, weight, out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.interpolation_for_backward
        b, c, n = grad_out.size()
        grad = _new(grad_out, (b, c, m), torch.float32, zero=True)
        K.interpolation_backward_cuda(b, c, n, m, grad_out.detach().contiguous(), idx, weight, grad)
        return grad, None, None


interpolation = Interpolation.apply


class Grouping(Function):
    """features (b,c,n), idx (b,m,nsample) -> (b,c,m,nsample).  Reference: pointops.py:121-150."""

    @staticmethod
    def forward(ctx, features, idx):
        assert features.is_contiguous() and idx.is_contiguous()
        b, c, n = features.size()
        _, m, nsample = idx.size()
        out = _new(features, (b, c, m, nsample), torch.float32)
        K.grouping_forward_cuda(b, c, n, m, nsample, features, idx, out)
        ctx.for_backwards = (idx, n)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, n = ctx.for_backwards
        b, c, m, nsample = grad_out.size()
        grad = _new(grad_out, (b, c, n), torch.float32, zero=True)
        K.grouping_backward_cuda(b, c, n, m, nsample, grad_out.detach().contiguous(), idx, grad)
        return grad, None


grouping = Grouping.apply


class GroupingInt(Function):
    """int64 payload variant of grouping.  Reference: pointops.py:153-172."""

    @staticmethod
    def forward(ctx, features, idx):
        assert features.is_contiguous() and idx.is_contiguous()
        b, c, n = features.size()
        _, m, nsample = idx.size()
        out = _new(features, (b, c, m, nsample), torch.int64)
        K.grouping_int_forward_cuda(b, c, n, m, nsample, features, idx, out)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None


grouping_int = GroupingInt.apply


class BallQuery(Function):
    """radius, nsample, xyz (b,n,3), new_xyz (b,m,3) -> idx (b,m,nsample).  Reference: pointops.py:175-198."""

    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        assert xyz.is_contiguous() and new_xyz.is_contiguous()
        b, n, _ = xyz.size()
        m = new_xyz.size(1)
        idx = _new(xyz, (b, m, nsample), torch.int32, zero=True)
        K.ballquery_cuda(b, n, m, radius, nsample, new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ballquery = BallQuery.apply


class FeatureDistribute(Function):
    """max_xyz (b,n,3), xyz (b,m,3) -> nearest-centre idx (b,m).  Reference: pointops.py:201-222."""

    @staticmethod
    def forward(ctx, max_xyz, xyz):
        assert max_xyz.is_contiguous() and xyz.is_contiguous()
        b, n, _ = max_xyz.size()
        m = xyz.size(1)
        out = _new(xyz, (b, m), torch.int32, zero=True)
        K.featuredistribute_cuda(b, n, m, max_xyz, xyz, out)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None


featuredistribute = FeatureDistribute.apply


class FeatureGather(Function):
    """max_feature (b,c,n), distribute_idx (b,m) -> (b,c,m).  Reference: pointops.py:225-257."""

    @staticmethod
    def forward(ctx, max_feature, distribute_idx):
        assert max_feature.is_contiguous() and distribute_idx.is_contiguous()
        b, c, n = max_feature.size()
        m = distribute_idx.size(1)
        out = _new(max_feature, (b, c, m), torch.float32, zero=True)
        K.featuregather_forward_cuda(b, n, m, c, max_feature, distribute_idx, out)
        ctx.for_backwards = (distribute_idx, n)
        return out

    @staticmethod
    def backward(ctx, grad):
        distribute_idx, n = ctx.for_backwards
        b, c, m = grad.size()
        out = _new(grad, (b, c, n), torch.float32, zero=True)
        K.featuregather_backward_cuda(b, n, m, c, grad.detach().contiguous(), distribute_idx, out)
        return out, None


featuregather = FeatureGather.apply


class LabelStatBallRange(Function):
    """Reference: pointops.py:260-286."""

    @staticmethod
    def forward(ctx, radius, xyz, new_xyz, label_stat):
        assert xyz.is_contiguous() and new_xyz.is_contiguous() and label_stat.is_contiguous()
        b, n, nclass = label_stat.size()
        m = new_xyz.size(1)
        out = _new(xyz, (b, m, nclass), torch.int32, zero=True)
        K.labelstat_ballrange_cuda(b, n, m, radius, nclass, new_xyz, xyz, label_stat, out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


labelstat_ballrange = LabelStatBallRange.apply


class LabelStatIdx(Function):
    """Reference: pointops.py:289-314."""

    @staticmethod
    def forward(ctx, nsample, label_stat, idx):
        assert label_stat.is_contiguous() and idx.is_contiguous()
        b, n, nclass = label_stat.size()
        m = idx.size(1)
        out = _new(idx, (b, m, nclass), torch.int32, zero=True)
        K.labelstat_idx_cuda(b, n, m, nsample, nclass, label_stat, idx, out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None


labelstat_idx = LabelStatIdx.apply


class LabelStatAndBallQuery(Function):
    """Reference: pointops.py:317-346."""

    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz, label_stat):
        assert xyz.is_contiguous() and new_xyz.is_contiguous() and label_stat.is_contiguous()
        b, n, nclass = label_stat.size()
        m = new_xyz.size(1)
        out = _new(xyz, (b, m, nclass), torch.int32, zero=True)
        idx = _new(xyz, (b, m, nsample), torch.int32, zero=True)
        K.labelstat_and_ballquery_cuda(b, n, m, radius, nsample, nclass, new_xyz, xyz, label_stat, idx, out)
        return out, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None, None, None


labelstat_and_ballquery = LabelStatAndBallQuery.apply


def pairwise_distances(x, y=None):
    """dist[i,j] = ||x[i]-y[j]||^2 via the expanded form, clamped at 0.  Reference: pointops.py:349-364."""
    x_norm = (x ** 2).sum(1).view(-1, 1)
    if y is None:
        y = x
        y_norm = x_norm.view(1, -1)
    else:
        y_norm = (y ** 2).sum(1).view(1, -1)
    dist = x_norm + y_norm - 2.0 * torch.mm(x, y.transpose(0, 1))
    return torch.clamp(dist, 0.0, float("inf"))


def _naive_sorted_idx(xyz, new_xyz):
    if new_xyz is None:
        new_xyz = xyz
    b, m, _ = new_xyz.size()
    n = xyz.size(1)
    diff = new_xyz.repeat(1, 1, n).view(b, m * n, 3) - xyz.repeat(1, m, 1).view(b, m * n, 3)
    return torch.sort(diff.pow(2).sum(dim=2).view(b, m, n), dim=2)[1]


class KNNQueryNaive(Function):
    """Pure-torch kNN by full sort.  Reference: pointops.py:367-401."""

    @staticmethod
    def forward(ctx, nsample, xyz, new_xyz=None):
        return _naive_sorted_idx(xyz, new_xyz)[:, :, 0:nsample].int()

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None


knnquery_naive = KNNQueryNaive.apply


class KNNQuery(Function):
    """nsample, xyz (b,n,3), new_xyz (b,m,3) -> idx (b,m,nsample) int32 (distances dropped).  Reference: pointops.py:404-430."""

    @staticmethod
    def forward(ctx, nsample, xyz, new_xyz=None):
        if new_xyz is None:
            new_xyz = xyz
        assert xyz.is_contiguous() and new_xyz.is_contiguous()
        b, m, _ = new_xyz.size()
        n = xyz.size(1)
        idx = _new(xyz, (b, m, nsample), torch.int32, zero=True)
        K.knnquery_cuda(b, n, m, nsample, xyz, new_xyz, idx, None)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None


knnquery = KNNQuery.apply


class KNNQueryExclude(Function):
    """Like knnquery_naive but skips the nearest (the query itself).  Reference: pointops.py:433-470."""

    @staticmethod
    def forward(ctx, nsample, xyz, new_xyz=None):
        return _naive_sorted_idx(xyz, new_xyz)[:, :, 1:nsample + 1].int()

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None


knnquery_exclude = KNNQueryExclude.apply


def _neighbour_idx(radius, nsample, xyz, new_xyz):
    return ballquery(radius, nsample, xyz, new_xyz) if radius is not None else knnquery(nsample, xyz, new_xyz)


class QueryAndGroup(nn.Module):
    """Ball / kNN grouping with centred xyz.  Reference: pointops.py:473-516."""

    def __init__(self, radius=None, nsample=32, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz=None, features=None, idx=None):
        if new_xyz is None:
            new_xyz = xyz
        if idx is None:
            idx = _neighbour_idx(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return grouped_xyz
        grouped_features = grouping(features, idx)
        return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features


class QueryAndGroup_Edge(nn.Module):
    """EdgeConv grouping: [xyz_j - xyz_i ; f_j - f_i].  Reference: pointops.py:519-582.

    With ``knn_dilation > 1`` the reference queries ``knn_dilation * nsample`` neighbours and then keeps
    ``candidates[:, :, torch.randperm(nsample)]`` — a random ORDER of the nearest ``nsample`` (pointops.py:551-555).
    The same CPU-generator draw is made here so RNG streams and returned ``idx`` match element for element.
    """

    def __init__(self, radius=None, nsample=32, knn_dilation=1, use_xyz=True, ret_gxyz=False, ret_sample_idx=False):
        super().__init__()
        self.radius, self.nsample, self.knn_dilation, self.use_xyz = radius, nsample, knn_dilation, use_xyz
        self.ret_gxyz = ret_gxyz
        self.ret_sample_idx = ret_sample_idx

    def neighbour_idx(self, xyz, new_xyz):
        if self.radius is not None:
            return ballquery(self.radius, self.nsample, xyz, new_xyz)
        if self.knn_dilation > 1:
            # the first nsample of a sorted (dilation*nsample)-NN list are the sorted nsample-NN
            nearest = knnquery(self.nsample, xyz, new_xyz)
            if torch.cuda.is_current_stream_capturing():
                # CUDA-graph capture (training.GraphedTrainStep): a host-to-device copy of a fresh permutation cannot be
                # captured, so the neighbour ORDER is the one drawn at the last eager call and stays frozen in the graph —
                # everything downstream (max over K, chamfer) is order-invariant
                perm_dev = getattr(self, "_perm_dev", None)
                if perm_dev is None or perm_dev.device != nearest.device:
                    raise RuntimeError("QueryAndGroup_Edge: run one eager forward on this device before capturing a CUDA graph")
                return nearest[:, :, perm_dev].contiguous()
            perm = torch.randperm(self.nsample)
            self._perm_dev = perm.to(nearest.device)
            return nearest[:, :, self._perm_dev].contiguous()
        return knnquery(self.nsample, xyz, new_xyz)

    def forward(self, xyz, new_xyz=None, features=None, center_features=None, idx=None):
        if new_xyz is None:
            new_xyz = xyz
        if idx is None:
            idx = self.neighbour_idx(xyz, new_xyz)
        o_grouped_xyz = grouping(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = o_grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            grouped_features = grouping(features, idx)
            if grouped_features.size(3) > 1:
                grouped_features = grouped_features - center_features.unsqueeze(-1)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        res = new_features
        if self.ret_gxyz:
            res = res, o_grouped_xyz
        if self.ret_sample_idx:
            res = res, idx
        return res


class QueryAndGroup_Edge_Split(nn.Module):
    """EdgeConv grouping that also returns the grouped xyz.  Reference: pointops.py:584-635."""

    def __init__(self, radius=None, nsample=32, use_xyz=True, ret_gxyz=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_gxyz = ret_gxyz

    def forward(self, xyz, new_xyz=None, features=None, center_features=None, idx=None):
        if new_xyz is None:
            new_xyz = xyz
        if idx is None:
            idx = _neighbour_idx(self.radius, self.nsample, xyz, new_xyz)
        o_grouped_xyz = grouping(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = o_grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            grouped_features = grouping(features, idx)
            if grouped_features.size(3) > 1:
                grouped_features = grouped_features - center_features.unsqueeze(-1)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        return (new_features, o_grouped_xyz) if self.ret_gxyz else (new_features, grouped_xyz)


class GroupAll(nn.Module):
    """All points form one group.  Reference: pointops.py:637-661."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped_features = features.unsqueeze(2)
        return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
