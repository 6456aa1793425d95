"""ctypes bindings of oracle/_ref/*.so — the reference's OWN CUDA kernels compiled from /root/reference for sm_100
(oracle/Makefile `ref`).  TEST INFRASTRUCTURE, GPU box only.  Inputs/outputs are torch CUDA tensors; the reference
launchers use the legacy default stream, so every call is followed by a device synchronise.

Launcher signatures: libs/pointops/src/*/*_cuda_kernel.h (extern "C"), plus the shims in oracle/ref_shim*.cu.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_K = None
_T = None


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_kernels.so"))


def torch_kernels_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_torchkernels.so"))


def _k():
    global _K
    if _K is None:
        _K = C.CDLL(os.path.join(_HERE, "_ref", "libref_kernels.so"))
    return _K


def _t():
    global _T
    if _T is None:
        _T = C.CDLL(os.path.join(_HERE, "_ref", "libref_torchkernels.so"))
    return _T


def _p(t):
    assert t.is_cuda and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def _sync():
    torch.cuda.synchronize()
    rc = _k().ref_sync()
    if rc != 0:
        raise RuntimeError(f"reference kernel failed: cudaError {rc}")


def furthestsampling(xyz, m):
    b, n, _ = xyz.shape
    idx = torch.zeros(b, m, dtype=torch.int32, device=xyz.device)
    temp = torch.full((b, n), 1e10, dtype=torch.float32, device=xyz.device)
    torch.cuda.synchronize()
    _k().furthestsampling_cuda_launcher(b, n, m, _p(xyz), _p(temp), _p(idx))
    _sync()
    return idx, temp


def knnquery(nsample, xyz, new_xyz):
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros(b, m, nsample, dtype=torch.int32, device=xyz.device)
    dist2 = torch.zeros(b, m, nsample, dtype=torch.float32, device=xyz.device)
    torch.cuda.synchronize()
    _k().knnquery_cuda_launcher(b, n, m, nsample, _p(xyz), _p(new_xyz), _p(idx), _p(dist2), C.c_void_p(0))
    _sync()
    return idx


def ballquery(radius, nsample, xyz, new_xyz):
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros(b, m, nsample, dtype=torch.int32, device=xyz.device)
    torch.cuda.synchronize()
    _k().ballquery_cuda_launcher_fast(b, n, m, C.c_float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx), C.c_void_p(0))
    _sync()
    return idx


def nearestneighbor(unknown, known):
    b, n, _ = unknown.shape
    m = known.shape[1]
    d2 = torch.zeros(b, n, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.zeros(b, n, 3, dtype=torch.int32, device=unknown.device)
    torch.cuda.synchronize()
    _k().nearestneighbor_cuda_launcher_fast(b, n, m, _p(unknown), _p(known), _p(d2), _p(idx))
    _sync()
    return d2, idx


def interpolation(points, idx, weight):
    b, c, m = points.shape
    n = idx.shape[1]
    out = torch.zeros(b, c, n, dtype=torch.float32, device=points.device)
    torch.cuda.synchronize()
    _k().interpolation_forward_cuda_launcher_fast(b, c, m, n, _p(points), _p(idx), _p(weight), _p(out))
    _sync()
    return out


def interpolation_backward(grad_out, idx, weight, m):
    b, c, n = grad_out.shape
    out = torch.zeros(b, c, m, dtype=torch.float32, device=grad_out.device)
    torch.cuda.synchronize()
    # the launcher's parameter NAMES are swapped (n,c) but positions are (b, c, n, m): interpolation_cuda.cpp:33
    _k().interpolation_backward_cuda_launcher(b, c, n, m, _p(grad_out), _p(idx), _p(weight), _p(out))
    _sync()
    return out


def grouping(points, idx):
    b, c, n = points.shape
    _, m, ns = idx.shape
    out = torch.zeros(b, c, m, ns, dtype=torch.float32, device=points.device)
    torch.cuda.synchronize()
    _k().grouping_forward_cuda_launcher_fast(b, c, n, m, ns, _p(points), _p(idx), _p(out))
    _sync()
    return out


def grouping_backward(grad_out, idx, n):
    b, c, m, ns = grad_out.shape
    out = torch.zeros(b, c, n, dtype=torch.float32, device=grad_out.device)
    torch.cuda.synchronize()
    _k().grouping_backward_cuda_launcher(b, c, n, m, ns, _p(grad_out), _p(idx), _p(out))
    _sync()
    return out


def gathering(points, idx):
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.zeros(b, c, m, dtype=torch.float32, device=points.device)
    torch.cuda.synchronize()
    _k().gathering_forward_cuda_launcher(b, c, n, m, _p(points), _p(idx), _p(out))
    _sync()
    return out


def gathering_backward(grad_out, idx, n):
    b, c, m = grad_out.shape
    out = torch.zeros(b, c, n, dtype=torch.float32, device=grad_out.device)
    torch.cuda.synchronize()
    _k().gathering_backward_cuda_launcher(b, c, n, m, _p(grad_out), _p(idx), _p(out))
    _sync()
    return out


def featuredistribute(max_xyz, xyz):
    b, n, _ = max_xyz.shape
    m = xyz.shape[1]
    out = torch.zeros(b, m, dtype=torch.int32, device=xyz.device)
    torch.cuda.synchronize()
    _k().featuredistribute_cuda_launcher(b, n, m, _p(max_xyz), _p(xyz), _p(out), C.c_void_p(0))
    _sync()
    return out


def labelstat_and_ballquery(radius, nsample, xyz, new_xyz, label_stat):
    b, n, nclass = label_stat.shape
    m = new_xyz.shape[1]
    out = torch.zeros(b, m, nclass, dtype=torch.int32, device=xyz.device)
    idx = torch.zeros(b, m, nsample, dtype=torch.int32, device=xyz.device)
    torch.cuda.synchronize()
    _k().labelstat_and_ballquery_cuda_launcher_fast(b, n, m, C.c_float(radius), nsample, nclass, _p(new_xyz), _p(xyz),
                                                    _p(label_stat), _p(idx), _p(out), C.c_void_p(0))
    _sync()
    return out, idx


def knn_cuda_raw(ref, query, k):
    """ref (dim,nr), query (dim,nq) -> dist (k,nq), ind (k,nq) int64 1-based; knn.cpp:23-56."""
    dim, nr = ref.shape
    nq = query.shape[1]
    dist = torch.empty(nr, nq, dtype=torch.float32, device=ref.device)
    ind = torch.empty(k, nq, dtype=torch.int64, device=ref.device)
    torch.cuda.synchronize()
    rc = _k().ref_knn_device(_p(ref), nr, _p(query), nq, dim, k, _p(dist), _p(ind))
    if rc != 0:
        raise RuntimeError(f"reference knn failed: cudaError {rc}")
    return dist[:k].contiguous(), ind


def chamfer_forward(xyz1, xyz2):
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = torch.zeros(B, n, device=xyz1.device); d2 = torch.zeros(B, m, device=xyz1.device)
    i1 = torch.zeros(B, n, dtype=torch.int32, device=xyz1.device); i2 = torch.zeros(B, m, dtype=torch.int32, device=xyz1.device)
    torch.cuda.synchronize()
    rc = _t().ref_chamfer_forward(B, n, _p(xyz1), m, _p(xyz2), _p(d1), _p(d2), _p(i1), _p(i2))
    assert rc == 0, rc
    return d1, d2, i1, i2


def chamfer_backward(xyz1, xyz2, idx1, idx2, g1, g2):
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    gx1 = torch.zeros_like(xyz1); gx2 = torch.zeros_like(xyz2)
    torch.cuda.synchronize()
    rc = _t().ref_chamfer_backward(B, n, _p(xyz1), m, _p(xyz2), _p(idx1), _p(idx2), _p(g1), _p(g2), _p(gx1), _p(gx2))
    assert rc == 0, rc
    return gx1, gx2


def emd_forward(xyz1, xyz2, eps, iters):
    """emdFunction.forward buffers (emd_module.py:40-53) + the reference kernels."""
    b, n, _ = xyz1.shape
    dev = xyz1.device
    i32 = dict(dtype=torch.int32, device=dev)
    dist = torch.zeros(b, n, device=dev)
    assignment = torch.zeros(b, n, **i32) - 1
    assignment_inv = torch.zeros(b, n, **i32) - 1
    price = torch.zeros(b, n, device=dev)
    bid = torch.zeros(b, n, **i32)
    bid_inc = torch.zeros(b, n, device=dev)
    max_inc = torch.zeros(b, n, device=dev)
    unass_idx = torch.zeros(b * n, **i32)
    max_idx = torch.zeros(b * n, **i32)
    unass_cnt = torch.zeros(512, **i32); unass_cnt_sum = torch.zeros(512, **i32); cnt_tmp = torch.zeros(512, **i32)
    torch.cuda.synchronize()
    rc = _t().ref_emd_forward(b, n, _p(xyz1), _p(xyz2), _p(dist), _p(assignment), _p(price), _p(assignment_inv), _p(bid),
                              _p(bid_inc), _p(max_inc), _p(unass_idx), _p(unass_cnt), _p(unass_cnt_sum), _p(cnt_tmp),
                              _p(max_idx), C.c_float(eps), iters)
    assert rc == 1, rc
    return dist, assignment
