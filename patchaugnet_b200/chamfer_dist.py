"""Chamfer distance — mirror of ``libs/chamfer_dist/__init__.py`` (+ the pybind module ``chamfer``,
chamfer_cuda.cpp:22-39) over libpatchaug_b200.so.

``forward(xyz1, xyz2) -> [dist1, dist2, idx1, idx2]`` and ``backward(xyz1, xyz2, idx1, idx2, g1, g2) -> [gx1, gx2]``
keep the pybind signatures; ``ChamferFunction`` / ``ChamferDistanceL1`` / ``L2`` / ``L2_split`` keep the module API.
Unlike the reference (no checks, chamfer.cu:159-164) inputs are validated and made contiguous.
"""
import torch

from . import _lib as L


def _prep(t):
    L.require_cuda(t)
    return t.contiguous().float()


def forward(xyz1, xyz2):
    xyz1, xyz2 = _prep(xyz1), _prep(xyz2)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist1 = torch.zeros(B, n, dtype=torch.float32, device=xyz1.device)
    dist2 = torch.zeros(B, m, dtype=torch.float32, device=xyz1.device)
    idx1 = torch.zeros(B, n, dtype=torch.int32, device=xyz1.device)
    idx2 = torch.zeros(B, m, dtype=torch.int32, device=xyz1.device)
    L.check(L.lib().pab_chamfer_forward(B, n, L.ptr(xyz1), m, L.ptr(xyz2), L.ptr(dist1), L.ptr(dist2), L.ptr(idx1),
                                        L.ptr(idx2), L.stream_ptr()), "chamfer.forward")
    return [dist1, dist2, idx1, idx2]


def backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2):
    xyz1, xyz2 = _prep(xyz1), _prep(xyz2)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1, g2 = _prep(grad_dist1), _prep(grad_dist2)
    gx1 = torch.zeros_like(xyz1)
    gx2 = torch.zeros_like(xyz2)
    L.check(L.lib().pab_chamfer_backward(B, n, L.ptr(xyz1), m, L.ptr(xyz2), L.ptr(idx1.contiguous()), L.ptr(idx2.contiguous()),
                                         L.ptr(g1), L.ptr(g2), L.ptr(gx1), L.ptr(gx2), L.stream_ptr()), "chamfer.backward")
    return [gx1, gx2]


class ChamferFunction(torch.autograd.Function):
    """Reference: libs/chamfer_dist/__init__.py:13-26."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, dist2, idx1, idx2 = forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        gx1, gx2 = backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2)
        return gx1, gx2


def _drop_zero_rows(xyz1, xyz2, ignore_zeros):
    if xyz1.size(0) == 1 and ignore_zeros:   # libs/chamfer_dist/__init__.py:36-41
        xyz1 = xyz1[torch.sum(xyz1, dim=2).ne(0)].unsqueeze(dim=0)
        xyz2 = xyz2[torch.sum(xyz2, dim=2).ne(0)].unsqueeze(dim=0)
    return xyz1, xyz2


class ChamferDistanceL2(torch.nn.Module):
    """mean(d1) + mean(d2).  Reference: __init__.py:29-44."""

    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def forward(self, xyz1, xyz2):
        dist1, dist2 = ChamferFunction.apply(*_drop_zero_rows(xyz1, xyz2, self.ignore_zeros))
        return torch.mean(dist1) + torch.mean(dist2)


class ChamferDistanceL2_split(torch.nn.Module):
    """(mean(d1), mean(d2)).  Reference: __init__.py:46-61."""

    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def forward(self, xyz1, xyz2):
        dist1, dist2 = ChamferFunction.apply(*_drop_zero_rows(xyz1, xyz2, self.ignore_zeros))
        return torch.mean(dist1), torch.mean(dist2)


class ChamferDistanceL1(torch.nn.Module):
    """(mean sqrt d1 + mean sqrt d2) / 2.  Reference: __init__.py:63-84."""

    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def forward(self, xyz1, xyz2):
        dist1, dist2 = ChamferFunction.apply(*_drop_zero_rows(xyz1, xyz2, self.ignore_zeros))
        return (torch.mean(torch.sqrt(dist1)) + torch.mean(torch.sqrt(dist2))) / 2
