"""GPU tier: pin BOTH the CPU oracle and the new kernels to the reference's OWN CUDA kernels, compiled from
/root/reference for sm_100 into oracle/_ref/ (oracle/Makefile `ref`; built in the container, shipped by gpurun).
This is the check that the fp32 contraction order and tie-break rules restated in oracle/pointops_oracle.c are the
ones the stock build really has on a B200.  Skipped (not failed) when oracle/_ref is absent."""
import numpy as np
import pytest
import torch

import util
from oracle import ops, refgpu
from patchaugnet_b200 import chamfer_dist, knn_cuda, pointops

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refgpu.available(), reason="oracle/_ref not built")]
DEV = "cuda"


def _clouds(b, n, seed=0, dup=False):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-1, 1, (b, n, 3)).astype(np.float32)
    if dup:
        xyz[:, n // 2:] = xyz[:, : n - n // 2]
    return xyz


def _g(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("n,m", [(4096, 1024), (1024, 128), (128, 16), (100, 37), (16, 16)])
def test_reference_fps_equals_oracle_and_new_kernel(n, m):
    for dup in (False, True):
        xyz = np.stack([util.synthetic_cloud(i, n).numpy() for i in range(3)]) if not dup else _clouds(3, n, n, True)
        ref_idx, ref_temp = refgpu.furthestsampling(_g(xyz), m)
        assert np.array_equal(ref_idx.cpu().numpy(), ops.furthestsampling(xyz, m)), "oracle != reference kernel"
        assert torch.equal(pointops.furthestsampling(_g(xyz), m), ref_idx), "new kernel != reference kernel"


@pytest.mark.parametrize("n,m,k", [(4096, 1024, 40), (1024, 128, 40), (128, 16, 40), (300, 50, 7)])
def test_reference_knnquery_equals_oracle_and_new_kernel(n, m, k):
    for dup in (False, True):
        xyz = _clouds(2, n, n + k, dup)
        q = np.ascontiguousarray(xyz[:, :: max(1, n // m)][:, :m])
        ref_idx = refgpu.knnquery(k, _g(xyz), _g(q))
        assert np.array_equal(ref_idx.cpu().numpy(), ops.knnquery(k, xyz, q))
        assert torch.equal(pointops.knnquery(k, _g(xyz), _g(q)), ref_idx)


def test_reference_three_nn_interp_group_gather():
    unknown, known = _clouds(2, 4096, 1, True), _clouds(2, 1024, 2, True)
    rd2, ridx = refgpu.nearestneighbor(_g(unknown), _g(known))
    od2, oidx = ops.nearestneighbor(unknown, known)
    assert np.array_equal(ridx.cpu().numpy(), oidx) and np.array_equal(rd2.cpu().numpy(), od2)
    dist, idx = pointops.nearestneighbor(_g(unknown), _g(known))
    assert torch.equal(idx, ridx) and torch.allclose(dist, torch.sqrt(rd2), rtol=1e-6, atol=0)
    rng = np.random.default_rng(3)
    feats = rng.normal(size=(2, 64, 1024)).astype(np.float32)
    w = rng.uniform(0, 1, (2, 4096, 3)).astype(np.float32)
    rout = refgpu.interpolation(_g(feats), ridx, _g(w))
    assert np.array_equal(rout.cpu().numpy(), ops.interpolation(feats, oidx, w)), "interpolation fma order"
    assert torch.equal(pointops.interpolation(_g(feats), ridx, _g(w)), rout)
    gidx = rng.integers(0, 1024, (2, 200, 20)).astype(np.int32)
    assert torch.equal(pointops.grouping(_g(feats), _g(gidx)), refgpu.grouping(_g(feats), _g(gidx)))
    cidx = rng.integers(0, 1024, (2, 300)).astype(np.int32)
    assert torch.equal(pointops.gathering(_g(feats), _g(cidx)), refgpu.gathering(_g(feats), _g(cidx)))
    g = rng.normal(size=(2, 64, 200, 20)).astype(np.float32)
    assert torch.allclose(refgpu.grouping_backward(_g(g), _g(gidx), 1024).cpu(), torch.from_numpy(ops.grouping_backward(g, gidx, 1024)), atol=1e-4)


def test_reference_ballquery_featuredistribute_labelstat():
    xyz = _clouds(2, 2000, 4)
    q = np.ascontiguousarray(xyz[:, ::13])
    rb = refgpu.ballquery(0.15, 12, _g(xyz), _g(q))
    assert np.array_equal(rb.cpu().numpy(), ops.ballquery(0.15, 12, xyz, q))
    assert torch.equal(pointops.ballquery(0.15, 12, _g(xyz), _g(q)), rb)
    centres = _clouds(2, 30, 5)
    rf = refgpu.featuredistribute(_g(centres), _g(xyz))
    assert np.array_equal(rf.cpu().numpy(), ops.featuredistribute(centres, xyz))
    assert torch.equal(pointops.featuredistribute(_g(centres), _g(xyz)), rf)
    ls = np.random.default_rng(6).integers(0, 3, (2, 2000, 5)).astype(np.int32)
    rs, ri = refgpu.labelstat_and_ballquery(0.15, 12, _g(xyz), _g(q), _g(ls))
    s, i = pointops.labelstat_and_ballquery(0.15, 12, _g(xyz), _g(q), _g(ls))
    assert torch.equal(s, rs) and torch.equal(i, ri)


def test_reference_knn_cuda():
    rng = np.random.default_rng(7)
    for dim, n, k in ((5, 1000, 40), (256, 600, 101), (3, 77, 9)):
        ref = rng.random((dim, n)).astype(np.float32)
        qry = rng.random((dim, n // 2)).astype(np.float32)
        rd, ri = refgpu.knn_cuda_raw(_g(ref), _g(qry), k)
        od, oi = ops.knn_cuda_raw(ref, qry, k)
        assert np.array_equal(ri.cpu().numpy(), oi), "oracle != reference KNN_CUDA"
        assert np.allclose(rd.cpu().numpy(), od, rtol=1e-6, atol=1e-7)
        nd, ni = knn_cuda._knn.knn(_g(ref), _g(qry), k)
        assert torch.equal(ni, ri) and torch.allclose(nd, rd, rtol=1e-6, atol=1e-7)


@pytest.mark.skipif(not refgpu.torch_kernels_available(), reason="oracle/_ref/libref_torchkernels.so not built")
def test_reference_chamfer_and_emd():
    rng = np.random.default_rng(8)
    a = rng.uniform(-1, 1, (500, 20, 3)).astype(np.float32)
    b = rng.uniform(-1, 1, (500, 20, 3)).astype(np.float32)
    b[:, 10:] = b[:, :10]
    rd1, rd2, ri1, ri2 = refgpu.chamfer_forward(_g(a), _g(b))
    od1, od2, oi1, oi2 = ops.chamfer_forward(a, b)
    assert np.array_equal(ri1.cpu().numpy(), oi1) and np.array_equal(ri2.cpu().numpy(), oi2)
    assert np.array_equal(rd1.cpu().numpy(), od1) and np.array_equal(rd2.cpu().numpy(), od2)
    d1, d2, i1, i2 = chamfer_dist.forward(_g(a), _g(b))
    assert torch.equal(d1, rd1) and torch.equal(d2, rd2) and torch.equal(i1, ri1) and torch.equal(i2, ri2)
    g1 = _g(rng.normal(size=(500, 20)).astype(np.float32)); g2 = _g(rng.normal(size=(500, 20)).astype(np.float32))
    rg1, rg2 = refgpu.chamfer_backward(_g(a), _g(b), ri1, ri2, g1, g2)
    ng1, ng2 = chamfer_dist.backward(_g(a), _g(b), i1, i2, g1, g2)
    assert torch.allclose(ng1, rg1, atol=1e-5) and torch.allclose(ng2, rg2, atol=1e-5)
    # large clouds (the tiled kernel)
    a2 = rng.uniform(-1, 1, (2, 3000, 3)).astype(np.float32); b2 = rng.uniform(-1, 1, (2, 2500, 3)).astype(np.float32)
    r = refgpu.chamfer_forward(_g(a2), _g(b2))
    n = chamfer_dist.forward(_g(a2), _g(b2))
    assert all(torch.equal(x, y) for x, y in zip(r, n))
    # EMD: reference kernels == C oracle == new kernel.  The auction is deterministic unless a GetMax decision has two
    # candidates within 1e-6 (the reference then keeps the last writer; the oracle counts such decisions per cloud):
    # every cloud without one must give identical assignments, and at least one such cloud is required per size.
    from patchaugnet_b200 import emd_module
    for n, eps, iters, seed in [(1024, 0.05, 200, 31), (1024, 0.01, 4000, 32), (4096, 0.02, 1024, 41), (4096, 0.05, 300, 43)]:
        r2 = np.random.default_rng(seed)
        x1 = r2.random((3, n, 3)).astype(np.float32); x2 = r2.random((3, n, 3)).astype(np.float32)
        odist, oasg, _, _, ties = ops.emd_forward(x1, x2, eps, iters)
        rdist, rasg = refgpu.emd_forward(_g(x1), _g(x2), eps, iters)
        ndist, nasg = emd_module.emdModule()(_g(x1), _g(x2), eps, iters)
        assert np.array_equal(nasg.cpu().numpy(), oasg) and np.abs(ndist.cpu().numpy() - odist).max() < 1e-4
        clean = ties == 0
        assert clean.any(), (n, eps, iters, ties)
        agree = (rasg.cpu().numpy() == oasg).mean(axis=1)
        assert (agree[clean] == 1.0).all(), (n, eps, iters, ties, agree)
        assert np.abs(rdist.cpu().numpy() - odist)[clean].max() < 1e-4
        # a cloud that met a tie follows a different (equally valid) auction path from there on: only its cost is comparable
        assert np.abs(np.sqrt(rdist.cpu().numpy()).mean(axis=1) - np.sqrt(odist).mean(axis=1)).max() < 5e-3
