"""CPU tier: the oracle against independent brute-force restatements, the reference's only known-answer property
(KNN_CUDA distances == sklearn KDTree to 3 decimals, libs/KNN_CUDA/tests/test_knn_cuda.py:32-47) and the golden
vectors generated from the reference's own nn.Module code (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

import util
from oracle import model, ops


def _clouds(b, n, seed=0, dup=False):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-1, 1, (b, n, 3)).astype(np.float32)
    if dup:
        xyz[:, n // 2:] = xyz[:, : n - n // 2]      # exact duplicates -> real ties
    return xyz


def test_opt_n_threads_matches_reference_values():
    # 1024/1024/128 for n = 4096/1024/128 and 256/64/16 for n = 256/64/16
    assert [ops.opt_n_threads(n) for n in (4096, 1024, 128, 256, 64, 16, 100, 1, 3)] == [1024, 1024, 128, 256, 64, 16, 64, 1, 2]


def test_fps_properties():
    xyz = _clouds(3, 300, 1)
    idx = ops.furthestsampling(xyz, 64)
    assert idx.shape == (3, 64) and (idx[:, 0] == 0).all()
    for bi in range(3):
        assert len(set(idx[bi].tolist())) == 64                      # distinct points without duplicates in the cloud
    # the selected set is maximin: every later pick is the farthest point from the earlier picks (fp64 check)
    for bi in range(3):
        chosen = [0]
        for j in range(1, 64):
            d = np.min(((xyz[bi][:, None, :].astype(np.float64) - xyz[bi][chosen][None]) ** 2).sum(-1), axis=1)
            assert d[idx[bi, j]] >= d.max() * (1 - 1e-6)
            chosen.append(int(idx[bi, j]))


def test_fps_tie_break_is_bit_reversed_thread_order():
    # 8 points, all equidistant from point 0 except duplicates: block size 8, ties resolved by the reduction tree
    xyz = np.zeros((1, 8, 3), np.float32)
    xyz[0, 1:] = [1, 0, 0]                     # points 1..7 identical -> every step is a 7-way (then fewer) tie
    idx = ops.furthestsampling(xyz, 4)
    # step 1: candidates tid 1..7 all d=1; tree: smallest bit-reversed tid wins -> bitrev3: 1->4,2->2,3->6,4->1,5->5,6->3,7->7 => tid 4
    assert idx[0, 1] == 4
    # afterwards every distance is 0: the all-equal tie goes to bit-reversed 0 -> tid 0
    assert idx[0, 2] == 0 and idx[0, 3] == 0


def test_knnquery_matches_stable_sort():
    for dup in (False, True):
        xyz = _clouds(2, 257, 2, dup)
        q = xyz[:, ::7].copy()
        idx, d2 = ops.knnquery(33, xyz, q, return_dist=True)
        for bi in range(2):
            dx = (q[bi][:, None, 0] - xyz[bi][None, :, 0]).astype(np.float32)
            dy = (q[bi][:, None, 1] - xyz[bi][None, :, 1]).astype(np.float32)
            dz = (q[bi][:, None, 2] - xyz[bi][None, :, 2]).astype(np.float32)
            # same fp32 operation order as the oracle: fma(dz,dz, fma(dx,dx, dy*dy)) emulated in float64 then rounded
            t = (dy * dy).astype(np.float32)
            t = (dx.astype(np.float64) * dx + t).astype(np.float32)
            d = (dz.astype(np.float64) * dz + t).astype(np.float32)
            ref = np.argsort(d, axis=1, kind="stable")[:, :33]
            assert (ref == idx[bi]).all()
            assert np.allclose(np.take_along_axis(d, ref, 1), d2[bi], rtol=0, atol=0)


def test_knnquery_more_neighbours_than_points_pads_with_zero_index():
    xyz = _clouds(1, 5, 3)
    idx, d2 = ops.knnquery(8, xyz, xyz, return_dist=True)
    assert (idx[0, :, 5:] == 0).all() and np.isinf(d2[0, :, 5:]).all()   # besti init 0, best init 1e40 (-> inf as float)
    with pytest.raises(ValueError):
        ops.knnquery(201, xyz, xyz)


def test_three_nn_and_interpolation():
    unknown, known = _clouds(2, 100, 4), _clouds(2, 17, 5)
    d2, idx = ops.nearestneighbor(unknown, known)
    full = ((unknown[:, :, None, :].astype(np.float64) - known[:, None, :, :]) ** 2).sum(-1)
    assert (np.argsort(full, axis=2, kind="stable")[:, :, :3] == idx).all()
    w = np.random.default_rng(0).uniform(0, 1, (2, 100, 3)).astype(np.float32)
    feats = np.random.default_rng(1).normal(size=(2, 6, 17)).astype(np.float32)
    out = ops.interpolation(feats, idx, w)
    ref = (np.take_along_axis(feats[:, :, None, :].repeat(100, 2), idx[:, None].astype(np.int64).repeat(6, 1), 3) * w[:, None]).sum(-1)
    assert np.allclose(out, ref, atol=1e-6)
    g = np.random.default_rng(2).normal(size=(2, 6, 100)).astype(np.float32)
    gb = ops.interpolation_backward(g, idx, w, 17)
    assert np.isclose((gb * feats).sum(), (g * out).sum(), rtol=1e-4)       # adjoint identity <J^T g, f> = <g, J f>


def test_ballquery_semantics():
    xyz = _clouds(1, 64, 6)
    q = xyz[:, :4].copy()
    idx = ops.ballquery(0.5, 8, xyz, q)
    d = ((q[0][:, None] - xyz[0][None]) ** 2).sum(-1)
    for i in range(4):
        hits = np.nonzero(d[i] < 0.25)[0][:8]
        exp = np.full(8, hits[0]); exp[: len(hits)] = hits
        assert (idx[0, i] == exp).all()
    assert (ops.ballquery(1e-6, 4, xyz, (q + 5).astype(np.float32)) == 0).all()      # no hit -> zeros from the caller


def test_gather_group_adjoints():
    rng = np.random.default_rng(7)
    f = rng.normal(size=(2, 5, 40)).astype(np.float32)
    idx = rng.integers(0, 40, (2, 9)).astype(np.int32)
    out = ops.gathering(f, idx)
    assert (out == np.take_along_axis(f, idx[:, None].astype(np.int64).repeat(5, 1), 2)).all()
    g = rng.normal(size=out.shape).astype(np.float32)
    assert np.isclose((ops.gathering_backward(g, idx, 40) * f).sum(), (g * out).sum(), rtol=1e-4)
    gi = rng.integers(0, 40, (2, 9, 4)).astype(np.int32)
    go = ops.grouping(f, gi)
    assert go.shape == (2, 5, 9, 4) and go[1, 3, 2, 1] == f[1, 3, gi[1, 2, 1]]
    gg = rng.normal(size=go.shape).astype(np.float32)
    assert np.isclose((ops.grouping_backward(gg, gi, 40) * f).sum(), (gg * go).sum(), rtol=1e-4)
    assert (ops.grouping_int(np.arange(2 * 5 * 40).reshape(2, 5, 40), gi)[0, 1, 2, 3] == 40 + gi[0, 2, 3])


def test_chamfer_first_minimum_and_gradient():
    rng = np.random.default_rng(8)
    a = rng.uniform(-1, 1, (3, 20, 3)).astype(np.float32)
    b = rng.uniform(-1, 1, (3, 20, 3)).astype(np.float32)
    b[:, 10:] = b[:, :10]                                          # duplicates: first minimum must win
    d1, d2, i1, i2 = ops.chamfer_forward(a, b)
    full = ((a[:, :, None].astype(np.float64) - b[:, None]) ** 2).sum(-1)
    assert (i1 == full.argmin(2)).all() and (i1 < 10).all()
    assert np.allclose(d1, full.min(2), atol=1e-6) and np.allclose(d2, full.min(1), atol=1e-6)
    g1 = rng.normal(size=d1.shape).astype(np.float32); g2 = rng.normal(size=d2.shape).astype(np.float32)
    gx1, gx2 = ops.chamfer_backward(a, b, i1, i2, g1, g2)
    # finite-difference check of sum(g1*d1 + g2*d2) w.r.t. one coordinate (indices frozen)
    eps = 1e-3
    a2 = a.copy(); a2[0, 3, 1] += eps
    e1, e2, _, _ = ops.chamfer_forward(a2, b)
    num = ((g1 * e1).sum() + (g2 * e2).sum() - (g1 * d1).sum() - (g2 * d2).sum()) / eps
    assert abs(num - gx1[0, 3, 1]) < 5e-2 * max(1.0, abs(num))


def test_emd_oracle_is_an_eps_optimal_assignment():
    """emd_oracle.c (restating emd_cuda.cu:95-226): a converged auction is a permutation whose total Euclidean cost is
    within n*eps of the optimal assignment (Bertsekas' bound), computed independently by scipy's Hungarian solver."""
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(31)
    a = rng.random((1, 1024, 3)).astype(np.float32)
    b = rng.random((1, 1024, 3)).astype(np.float32)
    eps = 0.01
    dist, asg, price, rounds, ties = ops.emd_forward(a, b, eps, 4000)
    assert rounds[0] < 4000 and sorted(asg[0].tolist()) == list(range(1024))          # converged: one-to-one
    cost = np.sqrt(((a[0][:, None, :] - b[0][None]) ** 2).sum(-1))
    ri, ci = linear_sum_assignment(cost)
    opt = cost[ri, ci].sum()
    got = np.sqrt(dist[0].astype(np.float64)).sum()
    assert opt - 1e-3 <= got <= opt + 1024 * eps
    assert np.allclose(dist[0], ((a[0] - b[0][asg[0]]) ** 2).sum(-1), atol=1e-6)          # CalcDist
    # identical clouds: every point's best object is itself at value 3.0
    d0, a0, _, r0, _ = ops.emd_forward(a, a, eps, 50)
    assert np.array_equal(a0[0], np.arange(1024)) and (d0 == 0).all() and r0[0] == 1
    # unconverged run: the last round lets every remaining bidder take its bid (emd_cuda.cu:203 `last ||`)
    d1, a1, _, r1, _ = ops.emd_forward(a, b, 0.05, 20)
    assert r1[0] == 20 and a1.min() >= 0 and len(set(a1[0].tolist())) < 1024
    with pytest.raises(ValueError):
        ops.emd_forward(a[:, :1000], b[:, :1000], eps, 10)                                # n % 1024 != 0 (emd_cuda.cu:246-249)


def test_knn_cuda_matches_kdtree_known_answer():
    # the reference's own test: distances vs sklearn KDTree(leaf_size=100), decimal=3 (test_knn_cuda.py:32-47)
    from sklearn.neighbors import KDTree
    rng = np.random.default_rng(9)
    for k, n in ((10, 100), (2, 11), (40, 301)):
        x = rng.random((2, n, 5)).astype(np.float32)
        D, I = ops.knn_cuda(x, x, k, transpose_mode=True)
        for bi in range(2):
            dist, ind = KDTree(x[bi], leaf_size=100).query(x[bi], k=k)
            np.testing.assert_almost_equal(D[bi], dist, decimal=3)
        assert I.dtype == np.int64 and I.min() >= 0 and (I[:, :, 0] == np.arange(n)[None]).all()
    # non-transposed layout: (bs, dim, n)
    xt = np.ascontiguousarray(x.transpose(0, 2, 1))
    D2, I2 = ops.knn_cuda(xt, xt, k)
    assert np.array_equal(D2.transpose(0, 2, 1), D) and np.array_equal(I2.transpose(0, 2, 1), I)


def test_featuredistribute_labelstat():
    rng = np.random.default_rng(10)
    centres, pts = _clouds(1, 6, 11), _clouds(1, 50, 12)
    di = ops.featuredistribute(centres, pts)
    assert (di[0] == ((pts[0][:, None] - centres[0][None]) ** 2).sum(-1).argmin(1)).all()
    mf = rng.normal(size=(1, 4, 6)).astype(np.float32)
    assert (ops.featuregather(mf, di)[0, 2] == mf[0, 2, di[0]]).all()
    ls = rng.integers(0, 3, (1, 50, 5)).astype(np.int32)
    q = pts[:, :7].copy()
    stat, idx = ops.labelstat_and_ballquery(0.6, 4, pts, q, ls)
    assert (idx == ops.ballquery(0.6, 4, pts, q)).all()
    assert (ops.labelstat_idx(4, ls, idx)[0, 0] == ls[0, idx[0, 0]].sum(0)).all()
    full = ops.labelstat_ballrange(0.6, pts, q, ls)
    d = ((q[0][:, None] - pts[0][None]) ** 2).sum(-1)
    assert (full[0, 2] == ls[0, d[2] < 0.36].sum(0)).all()


# ---- golden vectors from the reference's own Python modules -------------------------------------------------------

def test_state_dict_manifest_matches_reference():
    man = json.load(open(os.path.join(util.GOLDEN, "patchaugnet_state_dict.json")))
    net = util.build_network()
    sd = net.state_dict()
    assert list(sd.keys()) == list(man["state_dict"].keys())
    assert all(list(sd[k].shape) == man["state_dict"][k] for k in sd)
    assert sum(p.numel() for p in net.parameters()) == man["n_params"] == 13470308


def test_oracle_model_matches_reference_forward_golden():
    g = np.load(os.path.join(util.GOLDEN, "patchaugnet_ref_forward.npz"))
    net = util.build_network()
    x = util.golden_batch("patchaugnet")
    d = g["desc"]
    # the golden is discriminative: descriptors of different clouds differ by O(0.1) per element, so the 1e-4 contract
    # is 1e-3 of the input-dependent signal (round 1: 5e-3, VERDICT weak #1)
    assert len(d) == 8 and min(np.abs(d[i] - d[j]).max() for i in range(8) for j in range(i)) > 0.1
    out = model.patchaugnet_forward(net.state_dict(), util.PATCHAUGNET_CFG, x.numpy(), perms=list(g["perms"]))
    for i in range(3):
        assert np.array_equal(out["center_idx_origin"][i], g[f"center_idx{i}"])          # bit-exact indices
        # activations reach |x| ~ 9 with the calibrated statistics; the restated BatchNorm formula rounds differently
        assert np.allclose(out["fp_features"][i].numpy()[:, :, :8, 0], g[f"fp{i}_head"], atol=5e-5, rtol=1e-5)
    assert np.abs(out["desc"].numpy() - g["desc"]).max() < 5e-6
    # float64 dense path agrees with the fp32 reference within the 1e-4 contract of the north star
    out64 = model.patchaugnet_forward(net.state_dict(), util.PATCHAUGNET_CFG, x.numpy(), perms=list(g["perms"]), dtype=torch.float64)
    assert np.abs(out64["desc"].numpy() - g["desc"]).max() < 1e-5


def test_pptnet_state_dict_and_oracle_match_reference_golden():
    man = json.load(open(os.path.join(util.GOLDEN, "pptnet_state_dict.json")))
    net = util.build_pptnet()
    sd = net.state_dict()
    assert list(sd.keys()) == list(man["state_dict"].keys()) and len(sd) == 258
    assert all(list(sd[k].shape) == man["state_dict"][k] for k in sd)
    assert sum(p.numel() for p in net.parameters()) == man["n_params"] == 13389226
    sa = net.backbone.SA_modules[0].sas[0]
    assert sa.q_conv.weight is sa.k_conv.weight                                  # tied projection, pptnet.py:254
    g = np.load(os.path.join(util.GOLDEN, "pptnet_ref_forward.npz"))
    x = util.golden_batch("pptnet")
    d = g["desc"]
    assert len(d) == 8 and min(np.abs(d[i] - d[j]).max() for i in range(8) for j in range(i)) > 0.1
    out = model.pptnet_forward(sd, util.PPTNET_CFG, x.numpy())
    for i in range(4):
        assert np.array_equal(out["center_idx_origin"][i], g[f"center_idx{i}"])
    assert np.abs(out["desc"].numpy() - g["desc"]).max() < 5e-6


def test_pointnetvlad_cpu_plumbing_config():
    """BASELINE.json configs[0]: PointNetVLAD forward on one synthetic 4096-pt submap, CPU."""
    from patchaugnet_b200.pointnet_vlad import PointNetVlad
    torch.manual_seed(0)
    net = PointNetVlad(num_points=4096, global_feat=True, feature_transform=True, max_pool=False, output_dim=256).eval()
    assert len(net.state_dict()) == 78 and sum(p.numel() for p in net.parameters()) == 19779145
    with torch.no_grad():
        out = net(util.synthetic_batch(1, 4096, 0))
    assert out.shape == (1, 256) and torch.isfinite(out).all()


def test_pointnetvlad_matches_reference_module_golden():
    """Values, not just shapes: tests/golden/pointnetvlad_ref_forward.npz is the output of the reference's own
    PointNetVlad.py (make_golden.py make_pointnetvlad) for seeded weights with calibrated BatchNorm statistics."""
    from patchaugnet_b200.pointnet_vlad import PointNetVlad
    man = json.load(open(os.path.join(util.GOLDEN, "pointnetvlad_state_dict.json")))
    g = np.load(os.path.join(util.GOLDEN, "pointnetvlad_ref_forward.npz"))
    net = PointNetVlad(num_points=4096, global_feat=True, feature_transform=True, max_pool=False, output_dim=256)
    sd = net.state_dict()
    assert list(sd.keys()) == list(man["state_dict"].keys()) and all(list(sd[k].shape) == man["state_dict"][k] for k in sd)
    sd = util.fill_state_dict_raw(sd, seed=55)
    for k in g.files:
        if k.startswith("bn:"):
            sd[k[3:]] = torch.from_numpy(g[k])
    net.load_state_dict(sd)
    net.eval()
    with torch.no_grad():
        desc = net(util.golden_batch("patchaugnet")[[0, 2, 3, 4]]).numpy()
    d = g["desc"]
    assert min(np.abs(d[i] - d[j]).max() for i in range(4) for j in range(i)) > 0.1       # a discriminative golden
    assert np.abs(desc - d).max() < 1e-5


def test_pointnet_decoder_matches_reference_module_golden():
    """PointNetDecoder values (eval and train-mode BatchNorm) against the reference module's outputs."""
    g = np.load(os.path.join(util.GOLDEN, "decoder_ref.npz"))
    dec = util.build_network().decoder
    f = torch.from_numpy(g["feats"])
    with torch.no_grad():
        dec.eval()
        assert np.abs(dec(f).numpy() - g["out_eval"]).max() < 1e-5
        dec.train()
        assert np.abs(dec(f).numpy() - g["out_train"]).max() < 1e-5
        dec.eval()


def test_losses_match_plain_restatement():
    from patchaugnet_b200 import losses
    g = torch.Generator().manual_seed(5)
    q, pos, neg, other = (torch.randn(4, n, 16, generator=g) for n in (1, 2, 14, 1))
    got = losses.quadruplet_loss(q, pos, neg, other, 0.5, 0.2, lazy=True)
    # reference formula (losses/pointnetvlad_loss.py:53-105) written out with broadcasting
    positive = ((pos - q) ** 2).sum(2).max(1)[0][:, None]
    first = (0.5 + positive - ((neg - q) ** 2).sum(2)).clamp(min=0).max(1)[0].mean()
    second = (0.2 + positive - ((neg - other) ** 2).sum(2)).clamp(min=0).max(1)[0].mean()
    assert torch.allclose(got, first + second)
    assert torch.allclose(losses.triplet_loss(q, pos, neg, 0.5, lazy=False),
                          (0.5 + positive - ((neg - q) ** 2).sum(2)).clamp(min=0).sum(1).mean())
    a, b, c = [torch.randn(16, generator=g) for _ in range(3)], [torch.randn(16, generator=g) for _ in range(3)], \
              [torch.randn(16, generator=g) for _ in range(3)]
    want = torch.stack([(x - y + 1e-6).norm() ** 2 for x, y in zip(a, b)]).mean() + \
        torch.stack([(0.8 - (x - y + 1e-6).norm()).clamp(min=0) ** 2 for x, y in zip(a, c)]).mean()
    assert torch.allclose(losses.contrastive_loss(a, b, c, 0.8), want, atol=1e-5)


def test_prepare_oracle_reproduces_the_reference_normalisation():
    """oracle/prepare.py against tests/golden/prepare_ref.npz (the reference's normalize_point_cloud, make_prepare_golden.py)."""
    import os
    from oracle import prepare
    import util
    g = np.load(os.path.join(util.GOLDEN, "prepare_ref.npz"))
    for i in range(3):
        for zoom in (True, False):
            pc, meta = prepare.get_pc(g[f"raw{i}"], g[f"offset{i}"], normalize=True, zoom=zoom)
            assert np.array_equal(pc, g[f"pc{i}_zoom{int(zoom)}"])
            assert meta["scale"] == g[f"scale{i}_zoom{int(zoom)}"] and np.array_equal(meta["trans"], g[f"trans{i}_zoom{int(zoom)}"])
    pc, meta = prepare.get_pc(g["raw0"], g["offset0"], normalize=False)
    assert np.array_equal(pc, g["raw0"] - g["offset0"]) and meta["scale"] == 1.0
