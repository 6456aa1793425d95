"""GPU tier: every hand-written kernel, called through the C ABI (patchaugnet_b200.pointops_cuda -> libpatchaug_b200.so),
against the CPU oracle on the same seeded inputs — bit-exact for indices, exact or 1e-6 for copied/accumulated floats —
including tie-heavy clouds (duplicates, zero padding), ragged sizes and the reference's edge behaviours."""
import numpy as np
import pytest
import torch

import util
from oracle import ops
from patchaugnet_b200 import _lib as L
from patchaugnet_b200 import pointops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _clouds(b, n, seed=0, dup=False, zero_tail=0):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-1, 1, (b, n, 3)).astype(np.float32)
    if dup:
        xyz[:, n // 2:] = xyz[:, : n - n // 2]
    if zero_tail:
        xyz[:, n - zero_tail:] = 0
    return xyz


def _g(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("n,m", [(4096, 1024), (1024, 128), (128, 16), (16, 16), (100, 37), (5000, 64), (1, 1), (9000, 33)])
def test_fps_bit_exact(n, m):
    for kw in (dict(), dict(dup=True), dict(zero_tail=max(1, n // 50))):
        xyz = _clouds(3, n, seed=n + m, **kw)
        want = ops.furthestsampling(xyz, m)
        got = pointops.furthestsampling(_g(xyz), m)
        assert got.dtype == torch.int32 and torch.equal(got.cpu(), torch.from_numpy(want)), (n, m, kw)


def test_fps_thread_layout_does_not_change_result():
    xyz = _clouds(2, 4096, 5, dup=True)
    want = ops.furthestsampling(xyz, 256)
    for threads in (128, 256, 512, 1024):           # tie-break is defined by the reference's block size, not ours
        L.lib().pab_tune_fps_threads(threads)
        try:
            got = pointops.furthestsampling(_g(xyz), 256)
        finally:
            L.lib().pab_tune_fps_threads(0)
        assert torch.equal(got.cpu(), torch.from_numpy(want)), threads


@pytest.mark.parametrize("n,m,b", [(4096, 1024, 7), (2048, 300, 4), (8192, 200, 2), (4096, 4096, 2), (4096, 2, 3)])
def test_fps_pruned_sampler_is_bit_exact(n, m, b):
    """The pruned sampler (Morton chunks + box bounds, fps.cu) must give the full scan's indices and `temp` bit for
    bit: against the C oracle, against the full-scan kernel, for every packing of clouds per CTA, on uniform,
    duplicate-heavy, zero-padded and structured (planes / cylinders) clouds."""
    import util
    from patchaugnet_b200 import pointops_cuda as K
    clouds = [_clouds(b, n, seed=n + m), _clouds(b, n, seed=n + m + 1, dup=True), _clouds(b, n, seed=n + m + 2, zero_tail=n // 50),
              np.stack([util.place_visit(40 + i, 0, n) for i in range(b)])]
    for xyz in clouds:
        temp_want = np.full((b, n), 1e10, np.float32)
        want = ops.furthestsampling(xyz, m, temp_want)
        L.lib().pab_tune_fps_pruned(1)
        try:
            for cpc in (1, 2, 3):
                L.lib().pab_tune_fps_clouds_per_cta(cpc)
                t = torch.full((b, n), 1e10, device=DEV)
                idx = torch.zeros(b, m, dtype=torch.int32, device=DEV)
                K.furthestsampling_cuda(b, n, m, _g(xyz), t, idx)
                assert torch.equal(idx.cpu(), torch.from_numpy(want)), (n, m, cpc)
                assert torch.equal(t.cpu(), torch.from_numpy(temp_want)), (n, m, cpc)
        finally:
            L.lib().pab_tune_fps_pruned(0)
            L.lib().pab_tune_fps_clouds_per_cta(1)
        full = pointops.furthestsampling(_g(xyz), m)              # the default full-scan sampler
        assert torch.equal(full.cpu(), torch.from_numpy(want))


def test_fps_pruned_sampler_continues_from_a_given_temp():
    """`temp` is an input too (pointops.py:21 fills it with 1e10, the ABI takes any state): sampling on from the
    distances an earlier call left behind must match the oracle doing the same."""
    from patchaugnet_b200 import pointops_cuda as K
    xyz = _clouds(2, 4096, 77)
    temp = np.full((2, 4096), 1e10, np.float32)
    ops.furthestsampling(xyz, 50, temp)
    want = ops.furthestsampling(xyz, 40, temp.copy())
    for pruned in (1, 0):
        t = torch.from_numpy(temp.copy()).to(DEV)
        final = temp.copy()
        ops.furthestsampling(xyz, 40, final)
        idx = torch.zeros(2, 40, dtype=torch.int32, device=DEV)
        L.lib().pab_tune_fps_pruned(pruned)
        try:
            K.furthestsampling_cuda(2, 4096, 40, _g(xyz), t, idx)
        finally:
            L.lib().pab_tune_fps_pruned(0)
        assert torch.equal(idx.cpu(), torch.from_numpy(want)) and torch.equal(t.cpu(), torch.from_numpy(final))


def test_fps_temp_is_updated_like_the_reference():
    from patchaugnet_b200 import pointops_cuda as K
    xyz = _clouds(2, 300, 6)
    temp = np.full((2, 300), 1e10, np.float32)
    want = ops.furthestsampling(xyz, 20, temp)          # oracle updates its copy in place
    t = torch.full((2, 300), 1e10, device=DEV)
    idx = torch.zeros(2, 20, dtype=torch.int32, device=DEV)
    K.furthestsampling_cuda(2, 300, 20, _g(xyz), t, idx)
    assert torch.equal(idx.cpu(), torch.from_numpy(want)) and torch.equal(t.cpu(), torch.from_numpy(temp))


@pytest.mark.parametrize("n,m,k", [(4096, 1024, 40), (1024, 128, 20), (128, 16, 40), (300, 300, 1), (257, 31, 33), (50, 7, 64),
                                   (700, 40, 200), (5000, 64, 20), (5, 5, 8)])
def test_knnquery_bit_exact(n, m, k):
    for kw in (dict(), dict(dup=True)):
        xyz = _clouds(2, n, seed=n + k, **kw)
        q = np.ascontiguousarray(xyz[:, :: max(1, n // m)][:, :m])
        want = ops.knnquery(k, xyz, q)
        got = pointops.knnquery(k, _g(xyz), _g(q))
        assert torch.equal(got.cpu(), torch.from_numpy(want)), (n, m, k, kw)


@pytest.mark.parametrize("n,m,k", [(256, 64, 20), (257, 300, 33), (1000, 77, 1), (4096, 1024, 20), (4096, 200, 64), (8192, 50, 40),
                                   (5000, 64, 20)])
def test_knnquery_indexed_bit_exact(n, m, k):
    """Morton-chunk index + box-distance pruning returns exactly the brute-force answer (distances, ties to the lower index),
    also on clouds where half the points are exact duplicates, a block of points coincides at the origin, or the queries lie
    far outside the cloud."""
    lib = L.lib()
    b = 2
    for kw in (dict(), dict(dup=True), dict(zero_tail=n // 3)):
        xyz = _clouds(b, n, seed=3 * n + k, **kw)
        q = np.ascontiguousarray(xyz[:, :: max(1, n // m)][:, :m]) if m <= n else _clouds(b, m, seed=n + 1)
        q[:, -1] += 7.0                                                  # one query far away from every chunk box
        wi, wd = ops.knnquery(k, xyz, q, return_dist=True)
        gx, gq = _g(xyz), _g(q)
        index = torch.empty(lib.pab_knn_index_bytes(b, n), dtype=torch.uint8, device=DEV)
        idx = torch.full((b, q.shape[1], k), -7, dtype=torch.int32, device=DEV)
        d2 = torch.zeros(b, q.shape[1], k, device=DEV)
        L.check(lib.pab_knn_build_index(b, n, L.ptr(gx), L.ptr(index), L.stream_ptr()), "index")
        L.check(lib.pab_knnquery_indexed(b, n, q.shape[1], k, L.ptr(index), L.ptr(gq), L.ptr(idx), L.ptr(d2), L.stream_ptr()), "knn")
        torch.cuda.synchronize()
        assert torch.equal(idx.cpu(), torch.from_numpy(wi)), (n, m, k, kw)
        assert torch.equal(d2.cpu(), torch.from_numpy(wd)), (n, m, k, kw)
        # the drop-in entry point picks the same path on its own when there are enough queries
        assert torch.equal(pointops.knnquery(k, gx, gq).cpu(), torch.from_numpy(wi))


def test_knnquery_dist2_and_limits():
    from patchaugnet_b200 import pointops_cuda as K
    xyz = _clouds(1, 64, 7)
    idx = torch.zeros(1, 64, 5, dtype=torch.int32, device=DEV)
    d2 = torch.zeros(1, 64, 5, device=DEV)
    K.knnquery_cuda(1, 64, 64, 5, _g(xyz), _g(xyz), idx, d2)
    wi, wd = ops.knnquery(5, xyz, xyz, return_dist=True)
    assert torch.equal(idx.cpu(), torch.from_numpy(wi)) and torch.equal(d2.cpu(), torch.from_numpy(wd))
    with pytest.raises(ValueError):
        pointops.knnquery(201, _g(xyz), _g(xyz))        # the reference's fixed best[200] (knnquery_cuda_kernel.cu:21-22)


@pytest.mark.parametrize("n,m", [(4096, 1024), (1024, 128), (128, 16), (37, 3), (100, 2500)])
def test_three_nn_bit_exact_and_weights(n, m):
    for kw in (dict(), dict(dup=True)):
        unknown, known = _clouds(2, n, n, **kw), _clouds(2, m, m + 1, **kw)
        if kw:
            known[:, : min(n, m)] = unknown[:, : min(n, m)]          # zero distances -> the 1e-8 guard matters
        d2, idx = ops.nearestneighbor(unknown, known)
        dist, gi = pointops.nearestneighbor(_g(unknown), _g(known))
        assert torch.equal(gi.cpu(), torch.from_numpy(idx))
        raw_d2 = torch.empty(2, n, 3, device=DEV)
        raw_i = torch.empty(2, n, 3, dtype=torch.int32, device=DEV)
        from patchaugnet_b200 import pointops_cuda as K
        K.nearestneighbor_cuda(2, n, m, _g(unknown), _g(known), raw_d2, raw_i)
        assert torch.equal(raw_d2.cpu(), torch.from_numpy(d2)) and torch.equal(raw_i, gi)     # squared distances bit-exact
        assert torch.allclose(dist.cpu(), torch.sqrt(torch.from_numpy(d2)), rtol=1e-6, atol=0)  # pointops.py:76 sqrt in torch
        # fused weights = patch_aug_net.py:351-353 on the same distances
        wi = torch.empty(2, n, 3, dtype=torch.int32, device=DEV)
        w = torch.empty(2, n, 3, device=DEV)
        gu, gk = _g(unknown), _g(known)            # keep the device tensors alive across the raw C-ABI call
        L.check(L.lib().pab_three_nn_weights(2, n, m, L.ptr(gu), L.ptr(gk), L.ptr(wi), L.ptr(w), L.stream_ptr()), "3nn")
        torch.cuda.synchronize()
        r = 1.0 / (torch.sqrt(torch.from_numpy(d2)) + 1e-8)
        want_w = r / r.sum(2, keepdim=True)
        assert torch.equal(wi.cpu(), torch.from_numpy(idx))
        assert torch.allclose(w.cpu(), want_w, rtol=2e-6, atol=0)


@pytest.mark.parametrize("n,m", [(4096, 1024), (1000, 256), (300, 4096), (5000, 777), (8192, 2048)])
def test_three_nn_indexed_bit_exact(n, m):
    """3-NN against the Morton-chunk index of the known points (with and without an index over the unknown points):
    same indices and weights as the brute-force scan, including duplicate / coincident points."""
    lib = L.lib()
    b = 2
    for kw in (dict(), dict(dup=True), dict(zero_tail=m // 3)):
        unknown, known = _clouds(b, n, n + 5, **({} if "zero_tail" in kw else kw)), _clouds(b, m, m + 6, **kw)
        if kw:
            known[:, : min(n, m) // 2] = unknown[:, : min(n, m) // 2]
        unknown[:, -1] -= 9.0                                              # a query far outside the known cloud
        d2, want_i = ops.nearestneighbor(unknown, known)
        r = 1.0 / (torch.sqrt(torch.from_numpy(d2)) + 1e-8)
        want_w = r / r.sum(2, keepdim=True)
        gu, gk = _g(unknown), _g(known)
        kidx = torch.empty(lib.pab_knn_index_bytes(b, m), dtype=torch.uint8, device=DEV)
        L.check(lib.pab_knn_build_index(b, m, L.ptr(gk), L.ptr(kidx), L.stream_ptr()), "index")
        uidx = None
        if 256 <= n <= 8192:
            uidx = torch.empty(lib.pab_knn_index_bytes(b, n), dtype=torch.uint8, device=DEV)
            L.check(lib.pab_knn_build_index(b, n, L.ptr(gu), L.ptr(uidx), L.stream_ptr()), "index")
        for ui in ([None, uidx] if uidx is not None else [None]):
            gi = torch.full((b, n, 3), -3, dtype=torch.int32, device=DEV)
            w = torch.zeros(b, n, 3, device=DEV)
            L.check(lib.pab_three_nn_weights_indexed(b, n, m, L.ptr(gu), L.ptr(ui), L.ptr(kidx), L.ptr(gi), L.ptr(w), L.stream_ptr()), "3nn")
            torch.cuda.synchronize()
            assert torch.equal(gi.cpu(), torch.from_numpy(want_i)), (n, m, kw, ui is not None)
            assert torch.allclose(w.cpu(), want_w, rtol=2e-6, atol=0)


def test_gather_group_interp_forward_backward():
    rng = np.random.default_rng(11)
    b, c, n, m, k = 2, 19, 333, 47, 6
    f = rng.normal(size=(b, c, n)).astype(np.float32)
    idx = rng.integers(0, n, (b, m)).astype(np.int32)
    gidx = rng.integers(0, n, (b, m, k)).astype(np.int32)
    ft = _g(f).requires_grad_(True)
    out = pointops.gathering(ft, _g(idx))
    assert torch.equal(out.detach().cpu(), torch.from_numpy(ops.gathering(f, idx)))
    g = rng.normal(size=out.shape).astype(np.float32)
    out.backward(_g(g))
    assert torch.allclose(ft.grad.cpu(), torch.from_numpy(ops.gathering_backward(g, idx, n)), atol=1e-5)
    ft = _g(f).requires_grad_(True)
    out = pointops.grouping(ft, _g(gidx))
    assert torch.equal(out.detach().cpu(), torch.from_numpy(ops.grouping(f, gidx)))
    g = rng.normal(size=out.shape).astype(np.float32)
    out.backward(_g(g))
    assert torch.allclose(ft.grad.cpu(), torch.from_numpy(ops.grouping_backward(g, gidx, n)), atol=1e-5)
    li = rng.integers(0, 2 ** 40, (b, 3, n))
    assert torch.equal(pointops.grouping_int(_g(li), _g(gidx)).cpu(), torch.from_numpy(ops.grouping_int(li, gidx)))
    i3 = rng.integers(0, n, (b, m, 3)).astype(np.int32)
    w = rng.uniform(0, 1, (b, m, 3)).astype(np.float32)
    ft = _g(f).requires_grad_(True)
    out = pointops.interpolation(ft, _g(i3), _g(w))
    assert torch.equal(out.detach().cpu(), torch.from_numpy(ops.interpolation(f, i3, w)))     # same fma order -> bit-exact
    g = rng.normal(size=out.shape).astype(np.float32)
    out.backward(_g(g))
    assert torch.allclose(ft.grad.cpu(), torch.from_numpy(ops.interpolation_backward(g, i3, w, n)), atol=1e-5)


def test_ballquery_labelstat_featuredistribute():
    xyz = _clouds(2, 3000, 12)
    q = np.ascontiguousarray(xyz[:, ::29])
    for r, ns in ((0.2, 16), (0.05, 4), (1e-6, 3), (5.0, 32)):
        assert torch.equal(pointops.ballquery(r, ns, _g(xyz), _g(q)).cpu(), torch.from_numpy(ops.ballquery(r, ns, xyz, q)))
    far = (q + 10).astype(np.float32)
    assert (pointops.ballquery(0.1, 5, _g(xyz), _g(far)) == 0).all()
    rng = np.random.default_rng(13)
    ls = rng.integers(0, 4, (2, 3000, 7)).astype(np.int32)
    stat, idx = pointops.labelstat_and_ballquery(0.2, 16, _g(xyz), _g(q), _g(ls))
    ws, wi = ops.labelstat_and_ballquery(0.2, 16, xyz, q, ls)
    assert torch.equal(stat.cpu(), torch.from_numpy(ws)) and torch.equal(idx.cpu(), torch.from_numpy(wi))
    assert torch.equal(pointops.labelstat_idx(16, _g(ls), idx).cpu(), torch.from_numpy(ops.labelstat_idx(16, ls, wi)))
    assert torch.equal(pointops.labelstat_ballrange(0.2, _g(xyz), _g(q), _g(ls)).cpu(),
                       torch.from_numpy(ops.labelstat_ballrange(0.2, xyz, q, ls)))
    centres = _clouds(2, 40, 14)
    di = pointops.featuredistribute(_g(centres), _g(xyz))
    assert torch.equal(di.cpu(), torch.from_numpy(ops.featuredistribute(centres, xyz)))
    mf = rng.normal(size=(2, 9, 40)).astype(np.float32)
    mft = _g(mf).requires_grad_(True)
    out = pointops.featuregather(mft, di)
    assert torch.equal(out.detach().cpu(), torch.from_numpy(ops.featuregather(mf, di.cpu().numpy())))
    out.sum().backward()
    assert torch.allclose(mft.grad.cpu(), torch.from_numpy(ops.featuregather_backward(np.ones_like(out.detach().cpu().numpy()), di.cpu().numpy(), 40)))


def test_query_and_group_edge_consumes_rng_like_the_reference():
    xyz = _g(_clouds(2, 512, 15))
    new_xyz = xyz[:, :32].contiguous()
    feats = xyz.transpose(1, 2).contiguous()
    centre = feats[:, :, :32].contiguous()
    grouper = pointops.QueryAndGroup_Edge(None, 8, knn_dilation=2, use_xyz=True, ret_sample_idx=True)
    torch.manual_seed(3)
    nf, idx = grouper(xyz, new_xyz, feats, centre)
    torch.manual_seed(3)
    perm = torch.randperm(8)
    after = torch.rand(1)
    torch.manual_seed(3)
    grouper(xyz, new_xyz, feats, centre)
    assert torch.equal(torch.rand(1), after)                           # exactly one randperm(nsample) draw
    cand = torch.from_numpy(ops.knnquery(16, xyz.cpu().numpy(), new_xyz.cpu().numpy()))
    assert torch.equal(idx.cpu(), cand[:, :, perm])                    # pointops.py:553-555
    assert nf.shape == (2, 6, 32, 8)
    assert torch.allclose(nf[:, :3], nf[:, 3:])                        # features are xyz here: both halves are xyz_j - xyz_i
