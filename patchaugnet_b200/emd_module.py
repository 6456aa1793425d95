"""Approximate EMD (auction) — mirror of ``libs/emd_module/emd_module.py`` and the pybind module ``emd``
(emd.cpp:14-30) over libpatchaug_b200.so.

``forward(...)`` / ``backward(...)`` keep the 16- / 5-argument pybind signatures and return codes
(1 ok, 0 CUDA error, -1 bad shape); ``emdFunction`` / ``emdModule`` keep the module API.  Unlike the reference wrapper
(emd_module.py:55-56 ignores the code and silently returns dist = 0 for unsupported shapes), a bad return raises.
"""
import torch
from torch import nn
from torch.autograd import Function

from . import _lib as L


def forward(xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_increments, max_increments, unass_idx,
            unass_cnt, unass_cnt_sum, cnt_tmp, max_idx, eps, iters):
    L.require_cuda(xyz1, xyz2)
    b, n, _ = xyz1.shape
    if xyz2.shape[1] != n:
        return -1
    p = L.ptr
    return L.lib().pab_emd_forward(b, n, p(xyz1), p(xyz2), p(dist), p(assignment), p(price), p(assignment_inv), p(bid),
                                   p(bid_increments), p(max_increments), p(unass_idx), p(unass_cnt), p(unass_cnt_sum),
                                   p(cnt_tmp), p(max_idx), float(eps), int(iters), L.stream_ptr())


def backward(xyz1, xyz2, gradxyz, graddist, idx):
    b, n, _ = xyz1.shape
    p = L.ptr
    return L.lib().pab_emd_backward(b, n, p(xyz1), p(xyz2), p(gradxyz), p(graddist), p(idx), L.stream_ptr())


class emdFunction(Function):
    """Reference: emd_module.py:29-70."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, eps, iters):
        batchsize, n, _ = xyz1.size()
        _, m, _ = xyz2.size()
        assert n == m
        assert xyz1.size()[0] == xyz2.size()[0]
        assert batchsize <= 512
        L.require_cuda(xyz1, xyz2)
        dev = xyz1.device
        xyz1 = xyz1.contiguous().float()
        xyz2 = xyz2.contiguous().float()
        i32 = dict(dtype=torch.int32, device=dev)
        dist = torch.zeros(batchsize, n, device=dev)
        assignment = torch.full((batchsize, n), -1, **i32)
        assignment_inv = torch.full((batchsize, m), -1, **i32)
        price = torch.zeros(batchsize, m, device=dev)
        bid = torch.zeros(batchsize, n, **i32)
        bid_increments = torch.zeros(batchsize, n, device=dev)
        max_increments = torch.zeros(batchsize, m, device=dev)
        unass_idx = torch.zeros(batchsize * n, **i32)
        max_idx = torch.zeros(batchsize * m, **i32)
        unass_cnt = torch.zeros(512, **i32)
        unass_cnt_sum = torch.zeros(512, **i32)
        cnt_tmp = torch.zeros(512, **i32)
        rc = forward(xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_increments, max_increments, unass_idx,
                     unass_cnt, unass_cnt_sum, cnt_tmp, max_idx, eps, iters)
        if rc != 1:
            raise ValueError("emd: unsupported shape (n must be a multiple of 1024, batch <= 512)" if rc == -1
                             else "emd: CUDA error")
        ctx.save_for_backward(xyz1, xyz2, assignment)
        ctx.mark_non_differentiable(assignment)
        return dist, assignment

    @staticmethod
    def backward(ctx, graddist, gradidx):
        xyz1, xyz2, assignment = ctx.saved_tensors
        graddist = graddist.contiguous()
        gradxyz1 = torch.zeros_like(xyz1)
        gradxyz2 = torch.zeros_like(xyz2)
        if backward(xyz1, xyz2, gradxyz1, graddist, assignment) != 1:
            raise L.PabError("emd backward: CUDA error")
        return gradxyz1, gradxyz2, None, None


class emdModule(nn.Module):
    def forward(self, input1, input2, eps, iters):
        return emdFunction.apply(input1, input2, eps, iters)
