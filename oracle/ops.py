"""ctypes bindings of liboracle.so on numpy arrays (TEST INFRASTRUCTURE — see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("pointops_oracle.c", "emd_oracle.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def _l(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(C.POINTER(C.c_int64))


def opt_n_threads(n):
    return lib().ora_opt_n_threads(int(n))


def furthestsampling(xyz, m, temp=None):
    """xyz (b,n,3) f32 -> idx (b,m) i32.  pointops.py:11-29."""
    xyz, px = _f(xyz)
    b, n, _ = xyz.shape
    temp = np.full((b, n), 1e10, np.float32) if temp is None else temp
    temp, pt = _f(temp)
    idx = np.zeros((b, m), np.int32)
    lib().ora_furthestsampling(b, n, m, px, pt, idx.ctypes.data_as(C.POINTER(C.c_int)))
    return idx


def gathering(points, idx):
    points, pp = _f(points)
    idx, pi = _i(idx)
    b, c, n = points.shape
    m = idx.shape[1]
    out = np.empty((b, c, m), np.float32)
    lib().ora_gathering_forward(b, c, n, m, pp, pi, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def gathering_backward(grad_out, idx, n):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    b, c, m = grad_out.shape
    out = np.zeros((b, c, n), np.float32)
    lib().ora_gathering_backward(b, c, n, m, pg, pi, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def knnquery(nsample, xyz, new_xyz=None, return_dist=False):
    xyz, px = _f(xyz)
    new_xyz = xyz if new_xyz is None else new_xyz
    new_xyz, pq = _f(new_xyz)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = np.zeros((b, m, nsample), np.int32)
    d2 = np.zeros((b, m, nsample), np.float32)
    rc = lib().ora_knnquery(b, n, m, nsample, px, pq, idx.ctypes.data_as(C.POINTER(C.c_int)),
                            d2.ctypes.data_as(C.POINTER(C.c_float)))
    if rc != 0:
        raise ValueError("nsample must be <= 200 (knnquery_cuda_kernel.cu:21-22)")
    return (idx, d2) if return_dist else idx


def ballquery(radius, nsample, xyz, new_xyz):
    xyz, px = _f(xyz)
    new_xyz, pq = _f(new_xyz)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = np.zeros((b, m, nsample), np.int32)
    lib().ora_ballquery(b, n, m, C.c_float(radius), nsample, pq, px, idx.ctypes.data_as(C.POINTER(C.c_int)))
    return idx


def grouping(points, idx):
    points, pp = _f(points)
    idx, pi = _i(idx)
    b, c, n = points.shape
    _, m, ns = idx.shape
    out = np.empty((b, c, m, ns), np.float32)
    lib().ora_grouping_forward(b, c, n, m, ns, pp, pi, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def grouping_backward(grad_out, idx, n):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    b, c, m, ns = grad_out.shape
    out = np.zeros((b, c, n), np.float32)
    lib().ora_grouping_backward(b, c, n, m, ns, pg, pi, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def grouping_int(points, idx):
    points, pp = _l(points)
    idx, pi = _i(idx)
    b, c, n = points.shape
    _, m, ns = idx.shape
    out = np.empty((b, c, m, ns), np.int64)
    lib().ora_grouping_int_forward(b, c, n, m, ns, pp, pi, out.ctypes.data_as(C.POINTER(C.c_int64)))
    return out


def nearestneighbor(unknown, known):
    """-> (dist2 (b,n,3) SQUARED, idx (b,n,3)); pointops.py:76 takes the sqrt in Python."""
    unknown, pu = _f(unknown)
    known, pk = _f(known)
    b, n, _ = unknown.shape
    m = known.shape[1]
    d2 = np.empty((b, n, 3), np.float32)
    idx = np.empty((b, n, 3), np.int32)
    lib().ora_nearestneighbor(b, n, m, pu, pk, d2.ctypes.data_as(C.POINTER(C.c_float)),
                              idx.ctypes.data_as(C.POINTER(C.c_int)))
    return d2, idx


def interpolation(points, idx, weight):
    points, pp = _f(points)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    b, c, m = points.shape
    n = idx.shape[1]
    out = np.empty((b, c, n), np.float32)
    lib().ora_interpolation_forward(b, c, m, n, pp, pi, pw, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def interpolation_backward(grad_out, idx, weight, m):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    b, c, n = grad_out.shape
    out = np.zeros((b, c, m), np.float32)
    lib().ora_interpolation_backward(b, c, n, m, pg, pi, pw, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def featuredistribute(max_xyz, xyz):
    max_xyz, pm = _f(max_xyz)
    xyz, px = _f(xyz)
    b, n, _ = max_xyz.shape
    m = xyz.shape[1]
    out = np.zeros((b, m), np.int32)
    lib().ora_featuredistribute(b, n, m, pm, px, out.ctypes.data_as(C.POINTER(C.c_int)))
    return out


def featuregather(max_feature, distribute_idx):
    max_feature, pf = _f(max_feature)
    distribute_idx, pi = _i(distribute_idx)
    b, c, n = max_feature.shape
    m = distribute_idx.shape[1]
    out = np.zeros((b, c, m), np.float32)
    lib().ora_featuregather_forward(b, n, m, c, pf, pi, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def featuregather_backward(grad, distribute_idx, n):
    grad, pg = _f(grad)
    distribute_idx, pi = _i(distribute_idx)
    b, c, m = grad.shape
    out = np.zeros((b, c, n), np.float32)
    lib().ora_featuregather_backward(b, n, m, c, pg, pi, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def labelstat_idx(nsample, label_stat, idx):
    label_stat, pl = _i(label_stat)
    idx, pi = _i(idx)
    b, n, nclass = label_stat.shape
    m = idx.shape[1]
    out = np.zeros((b, m, nclass), np.int32)
    lib().ora_labelstat_idx(b, n, m, nsample, nclass, pl, pi, out.ctypes.data_as(C.POINTER(C.c_int)))
    return out


def labelstat_ballrange(radius, xyz, new_xyz, label_stat):
    xyz, px = _f(xyz)
    new_xyz, pq = _f(new_xyz)
    label_stat, pl = _i(label_stat)
    b, n, nclass = label_stat.shape
    m = new_xyz.shape[1]
    out = np.zeros((b, m, nclass), np.int32)
    lib().ora_labelstat_ballrange(b, n, m, C.c_float(radius), nclass, pq, px, pl, out.ctypes.data_as(C.POINTER(C.c_int)))
    return out


def labelstat_and_ballquery(radius, nsample, xyz, new_xyz, label_stat):
    xyz, px = _f(xyz)
    new_xyz, pq = _f(new_xyz)
    label_stat, pl = _i(label_stat)
    b, n, nclass = label_stat.shape
    m = new_xyz.shape[1]
    out = np.zeros((b, m, nclass), np.int32)
    idx = np.zeros((b, m, nsample), np.int32)
    lib().ora_labelstat_and_ballquery(b, n, m, C.c_float(radius), nsample, nclass, pq, px, pl,
                                      idx.ctypes.data_as(C.POINTER(C.c_int)), out.ctypes.data_as(C.POINTER(C.c_int)))
    return out, idx


def chamfer_forward(xyz1, xyz2):
    """-> dist1 (B,n), dist2 (B,m), idx1, idx2 — chamfer_cuda.cpp:22-29 / chamfer.cu:147-171."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.zeros((B, n), np.float32); i1 = np.zeros((B, n), np.int32)
    d2 = np.zeros((B, m), np.float32); i2 = np.zeros((B, m), np.int32)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    lib().ora_chamfer_one_direction(B, n, p1, m, p2, d1.ctypes.data_as(fp), i1.ctypes.data_as(ip))
    lib().ora_chamfer_one_direction(B, m, p2, n, p1, d2.ctypes.data_as(fp), i2.ctypes.data_as(ip))
    return d1, d2, i1, i2


def chamfer_backward(xyz1, xyz2, idx1, idx2, g1, g2):
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    idx1, pi1 = _i(idx1)
    idx2, pi2 = _i(idx2)
    g1, pg1 = _f(g1)
    g2, pg2 = _f(g2)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    gx1 = np.zeros_like(xyz1); gx2 = np.zeros_like(xyz2)
    fp = C.POINTER(C.c_float)
    lib().ora_chamfer_grad_one_direction(B, n, p1, m, p2, pg1, pi1, gx1.ctypes.data_as(fp), gx2.ctypes.data_as(fp))
    lib().ora_chamfer_grad_one_direction(B, m, p2, n, p1, pg2, pi2, gx2.ctypes.data_as(fp), gx1.ctypes.data_as(fp))
    return gx1, gx2


def emd_forward(xyz1, xyz2, eps, iters):
    """emdFunction.forward (emd_module.py:29-70) over emd_oracle.c -> dist (b,n) f32, assignment (b,n) i32,
    price (b,n) f32, rounds used per cloud, number of GetMax decisions with more than one candidate per cloud."""
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    b, n, _ = xyz1.shape
    assert xyz2.shape == xyz1.shape
    dist = np.zeros((b, n), np.float32)
    price = np.zeros((b, n), np.float32)
    asg = np.zeros((b, n), np.int32)
    rounds = np.zeros(b, np.int32)
    ties = np.zeros(b, np.int32)
    ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
    rc = lib().ora_emd_forward(b, n, p1, p2, C.c_float(eps), int(iters), dist.ctypes.data_as(fp), asg.ctypes.data_as(ip),
                               price.ctypes.data_as(fp), rounds.ctypes.data_as(ip), ties.ctypes.data_as(ip))
    if rc != 1:
        raise ValueError("emd oracle: unsupported shape")
    return dist, asg, price, rounds, ties


def knn_cuda_raw(ref, query, k):
    """ref (dim,nr), query (dim,nq) -> dist (k,nq) f32 (sqrt), ind (k,nq) int64 1-based.  knn.cpp:23-56."""
    ref, pr = _f(ref)
    query, pq = _f(query)
    dim, nr = ref.shape
    nq = query.shape[1]
    d = np.empty((k, nq), np.float32)
    i = np.empty((k, nq), np.int64)
    rc = lib().ora_knn_cuda(pr, nr, pq, nq, dim, k, d.ctypes.data_as(C.POINTER(C.c_float)),
                            i.ctypes.data_as(C.POINTER(C.c_int64)))
    if rc != 0:
        raise ValueError("k must be in [1, nr]")
    return d, i


def knn_cuda(ref, query, k, transpose_mode=False):
    """KNN(k, transpose_mode).forward — knn_cuda/__init__.py:48-74.  Returns (D, I) with I 0-based int64."""
    D, I = [], []
    for r, q in zip(ref, query):
        if transpose_mode:
            r, q = r.T, q.T
        d, i = knn_cuda_raw(r, q, k)
        i = i - 1
        if transpose_mode:
            d, i = d.T, i.T
        D.append(np.ascontiguousarray(d)); I.append(np.ascontiguousarray(i))
    return np.stack(D), np.stack(I)
