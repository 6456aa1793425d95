"""GPU tier: the fused PatchAugNet descriptor path against the oracle forward, the golden vectors produced by the
reference's own nn.Module code, the op-by-op (reference-shaped) path on the same kernels, and size-independent
properties at the BASELINE.json workload size (B=32 x 4096 points)."""
import os

import numpy as np
import pytest
import torch

import util
from oracle import model
from patchaugnet_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-4        # north star: descriptors within 1e-4 (fp32) of the reference forward


@pytest.fixture(scope="module")
def net():
    return util.build_network(DEV)


def test_fused_forward_matches_reference_golden(net):
    g = np.load(os.path.join(util.GOLDEN, "patchaugnet_ref_forward.npz"))
    x = util.golden_batch("patchaugnet").to(DEV)          # 8 clouds: uniform, tie-stress, structured places
    before = L.lib().pab_num_launches()
    with torch.no_grad():
        desc, fp_features, center_idx = net(x)
    assert L.lib().pab_num_launches() - before == net.engine().launches_per_forward()   # the CUDA library did the work
    for i in range(3):
        assert torch.equal(center_idx[i].cpu(), torch.from_numpy(g[f"center_idx{i}"]))   # FPS indices bit-exact
    assert np.abs(desc.cpu().numpy() - g["desc"]).max() < TOL
    for i, f in enumerate(fp_features):
        assert tuple(f.shape) == (8, 256, (128, 1024, 4096)[i], 1)
        # features reach |x| ~ 9 with the calibrated statistics: 2e-4 of the largest magnitude (the reference's own fp32
        # features sit 3e-5 from a float64 evaluation); the contract quantity is the descriptor above
        assert np.abs(f[:, :, :8, 0].cpu().numpy() - g[f"fp{i}_head"]).max() < 2e-4 * max(1.0, np.abs(g[f"fp{i}_head"]).max())
        # per-channel sums over all points: mean absolute error below 2e-5 of the largest magnitude
        n_pts = f.shape[2]
        assert np.abs(f.sum(dim=(2, 3)).double().cpu().numpy() - g[f"fp{i}_sum"]).max() < 2e-5 * n_pts * max(1.0, np.abs(g[f"fp{i}_head"]).max())


def test_fused_forward_matches_oracle_on_fresh_inputs(net):
    x = util.synthetic_batch(3, 4096, start=50)
    out = model.patchaugnet_forward(net.state_dict(), util.PATCHAUGNET_CFG, x.numpy(), perms=[np.arange(20)] * 3)
    with torch.no_grad():
        desc, fp_features, center_idx = net(x.to(DEV))
    for i in range(3):
        assert np.array_equal(center_idx[i].cpu().numpy(), out["center_idx_origin"][i])
        assert np.abs(fp_features[i].cpu().numpy() - out["fp_features"][i].numpy()).max() < 5e-4 * max(1.0, out["fp_features"][i].abs().max().item())
    assert np.abs(desc.cpu().numpy() - out["desc"].numpy()).max() < TOL
    out64 = model.patchaugnet_forward(net.state_dict(), util.PATCHAUGNET_CFG, x.numpy(), perms=[np.arange(20)] * 3, dtype=torch.float64)
    assert np.abs(desc.cpu().numpy() - out64["desc"].numpy()).max() < TOL


def test_batch32_matches_oracle(net):
    """BASELINE.json config 2 at its full batch: 32 clouds (uniform, tie-stress and structured places mixed) through the
    fused path against the CPU oracle — indices bit-exact, descriptors within 1e-4."""
    x = torch.cat([util.synthetic_batch(14, 4096, start=600), util.tie_stress_cloud(5)[None, None],
                   util.tie_stress_cloud(6)[None, None], util.place_batch(range(100, 116), 0)], 0)
    assert x.shape[0] == 32
    out = model.patchaugnet_forward(net.state_dict(), util.PATCHAUGNET_CFG, x.numpy(), perms=[np.arange(20)] * 3)
    with torch.no_grad():
        desc, fp_features, center_idx = net(x.to(DEV))
    for i in range(3):
        assert np.array_equal(center_idx[i].cpu().numpy(), out["center_idx_origin"][i])
    want = out["desc"].numpy()
    assert min(np.abs(want[i] - want[j]).max() for i in range(32) for j in range(i)) > 0.05    # not a collapsed descriptor
    assert np.abs(desc.cpu().numpy() - want).max() < TOL
    for i in range(3):
        ref = out["fp_features"][i].numpy()
        assert np.abs(fp_features[i].cpu().numpy() - ref).max() < 5e-4 * max(1.0, np.abs(ref).max())


def test_op_by_op_path_matches_fused_path(net):
    x = util.synthetic_batch(2, 4096, start=80).to(DEV)
    with torch.no_grad():
        d_fused, f_fused, c_fused = net(x)
        net.use_fused = False
        try:
            d_ops, f_ops, c_ops = net(x)
        finally:
            net.use_fused = True
    for a, b in zip(c_fused, c_ops):
        assert torch.equal(a, b)
    assert (d_fused - d_ops).abs().max().item() < TOL
    for a, b in zip(f_fused, f_ops):
        assert a.shape == b.shape and (a - b).abs().max().item() < 5e-4 * max(1.0, b.abs().max().item())


def test_training_path_backward_runs(net):
    net.train()
    try:
        x = util.synthetic_batch(2, 1024, start=90).to(DEV).requires_grad_(True)
        cfg = dict(util.PATCHAUGNET_CFG, SAMPLING=[256, 64, 16], MAX_SAMPLES=[64, 256, 1024])
        small = util.build_network(DEV, cfg=cfg).train()
        nn_dict = {(0, 1): [[0, 1]]}
        (desc, recon), fp_features, center_idx = small(x, nn_dict)
        from patchaugnet_b200.chamfer_dist import ChamferDistanceL1
        loss = desc.pow(2).sum() + ChamferDistanceL1()(torch.cat(recon["origin_patches"]), torch.cat(recon["reconstructed_patches"]))
        loss.backward()
        grads = [p.grad for p in small.parameters() if p.grad is not None]
        assert len(grads) > 50 and all(torch.isfinite(g).all() for g in grads)
        assert x.grad is not None and torch.isfinite(x.grad).all()
    finally:
        net.eval()


def test_full_size_properties(net):
    """BASELINE.json config 2: batch 32 x 4096 points."""
    x = util.synthetic_batch(32, 4096, start=100).to(DEV)
    with torch.no_grad():
        desc, fp_features, center_idx = net(x)
        assert desc.shape == (32, 256) and torch.isfinite(desc).all()
        assert torch.allclose(desc.norm(dim=1), torch.ones(32, device=DEV), atol=1e-5)           # F.normalize at the end
        # every cloud is an independent unit: reversing the batch reverses the outputs, bit for bit
        d2, f2, c2 = net(x.flip(0))
        assert torch.equal(d2.flip(0), desc) and torch.equal(c2[0].flip(0), center_idx[0])
        # determinism
        d3, _, _ = net(x)
        assert torch.equal(d3, desc)
        # FPS picks distinct points, first pick is index 0, deeper levels index into the original cloud
        c0 = center_idx[0].cpu().numpy()
        assert (c0[:, 0] == 0).all() and all(len(set(r.tolist())) == 1024 for r in c0)
        assert all(set(center_idx[1][i].tolist()) <= set(center_idx[0][i].tolist()) for i in range(32))
        # single-cloud batches give the same descriptors as the big batch
        d1, _, _ = net(x[5:6])
        assert torch.equal(d1[0], desc[5])


def test_cuda_graph_replay_matches_eager(net):
    x = util.synthetic_batch(4, 4096, start=200).to(DEV)
    with torch.no_grad():
        eager, _, ce = net(x)
        eng = net.engine()
        eng.capture_graph(4, 4096)
        try:
            replay, _, cr = net(x)
            y = util.synthetic_batch(4, 4096, start=300).to(DEV)
            r2, _, _ = net(y)
        finally:
            eng._graphs.clear()
        e2, _, _ = net(y)
    assert torch.equal(eager, replay) and torch.equal(ce[2], cr[2]) and torch.equal(r2, e2)


def test_state_dict_roundtrip_keeps_reference_layout(net):
    sd = net.state_dict()
    assert "backbone.SA_modules.0.mlps.0.layer0.conv.weight" in sd and "aggregation.afa.fc.weight" in sd
    assert "aggregation.vlads.2.hidden1_weights" in sd and "aggregation.afa.mlpa.trans_conv.weight" in sd
    ckpt = {"state_dict_encoder": sd}                       # train_place_recognition.py:183-184 checkpoint layout
    other = util.build_network(DEV, seed=999)
    x = util.synthetic_batch(1, 4096, 7).to(DEV)
    with torch.no_grad():
        a, _, _ = other(x)
        other.load_state_dict(ckpt["state_dict_encoder"])
        other.eval()
        b, _, _ = other(x)
        c, _, _ = net(x)
    assert not torch.equal(a, b) and torch.equal(b, c)      # engine refolds after load_state_dict


def test_pipelined_forward_stream_matches_per_batch_forward(net):
    """Throughput mode (two streams, ping-pong workspaces) must give exactly the per-batch results, in order."""
    batches = [util.synthetic_batch(4, 4096, start=400 + 4 * i).to(DEV) for i in range(5)]
    with torch.no_grad():
        want = torch.cat([net(b, return_feat=False) for b in batches])
        got = net.engine().forward_stream(batches)
        torch.cuda.synchronize()
        assert torch.equal(got, want)
        # through the retrieval API with host clouds and a ragged tail batch
        from patchaugnet_b200 import retrieval
        host = torch.cat([b.squeeze(1) for b in batches]).cpu()[:18].pin_memory()
        d = retrieval.extract_descriptors(net, host, batch_size=4, device=torch.device(DEV))
        torch.cuda.synchronize()
    assert torch.equal(d, want[:18])
    # database and query sets through ONE pipelined sequence (retrieval.extract_descriptor_sets): same descriptors, set by set
    with torch.no_grad():
        a, b = retrieval.extract_descriptor_sets(net, [host[:13], host[5:18]], batch_size=4, device=torch.device(DEV))
        torch.cuda.synchronize()
    assert torch.equal(a, want[:13]) and torch.equal(b, want[5:18])
    # consecutive batches concatenated into larger launch sequences (coalesce): submaps are independent and every kernel's arithmetic
    # depends on the cloud only, so the descriptors must not change by a bit — groups of two batches plus an uncoalesced remainder
    with torch.no_grad():
        got2 = net.engine().forward_stream(batches, coalesce=8)
        d2 = retrieval.extract_descriptors(net, host, batch_size=4, device=torch.device(DEV), launch_batch=8)
        torch.cuda.synchronize()
    assert torch.equal(got2, want) and torch.equal(d2, want[:18])


@pytest.mark.parametrize("agg_type,gating", [(0, False), (1, False), (3, False), (4, False), (5, False), (2, True), (0, True)])
def test_other_aggregation_variants_through_the_fused_engine(agg_type, gating):
    """aggregation_type 0/1/3/4/5 and GATING (reference loupe.py:289-328) — fused backbone + NetVLAD
    kernels, the variant's own tail — against the op-by-op path."""
    cfg = dict(util.PATCHAUGNET_CFG, SAMPLING=[256, 64, 16], MAX_SAMPLES=[64, 256, 1024], AGGREGATION_TYPE=agg_type, GATING=gating)
    small = util.build_network(DEV, cfg=cfg)
    assert small.fusable() and small.engine().fused_tail == (agg_type == 2 and not gating)
    x = util.synthetic_batch(3, 1024, start=300).to(DEV)
    before = L.lib().pab_num_launches()
    with torch.no_grad():
        d_fused, f_fused, c_fused = small(x)
        assert L.lib().pab_num_launches() - before == small.engine().launches_per_forward()
        small.use_fused = False
        d_ops, f_ops, c_ops = small(x)
    assert d_fused.shape == d_ops.shape and torch.isfinite(d_fused).all()
    assert (d_fused - d_ops).abs().max().item() < TOL
    for a, b in zip(c_fused, c_ops):
        assert torch.equal(a, b)


def test_ball_query_grouper_through_the_fused_engine():
    """QueryAndGroup_Edge with a radius (pointops.py:548-549) inside the fused SA modules."""
    cfg = dict(util.PATCHAUGNET_CFG, SAMPLING=[256, 64, 16], MAX_SAMPLES=[64, 256, 1024])
    small = util.build_network(DEV, cfg=cfg)
    for mod, r in zip(small.backbone.SA_modules, (0.25, 0.5, 1.0)):
        mod.groupers[0].radius = r
    small._engine = None
    x = util.synthetic_batch(2, 1024, start=310).to(DEV)
    with torch.no_grad():
        d_fused, f_fused, c_fused = small(x)
        small.use_fused = False
        d_ops, f_ops, c_ops = small(x)
        small.use_fused = True
        for mod in small.backbone.SA_modules:
            mod.groupers[0].radius = None
        small._engine = None
        d_knn, _, _ = small(x)
    assert (d_fused - d_ops).abs().max().item() < TOL
    assert (d_fused - d_knn).abs().max().item() > 1e-3          # the grouper really changed
    for a, b in zip(f_fused, f_ops):
        assert (a - b).abs().max().item() < 5e-4 * max(1.0, b.abs().max().item())


def test_single_neighbour_groups_take_the_op_path_and_match_the_oracle_rule():
    """QueryAndGroup_Edge skips the centre-feature subtraction when nsample == 1 (pointops.py:562-563); the fused
    loaders always subtract, so such a configuration must not be fused."""
    cfg = dict(util.PATCHAUGNET_CFG, SAMPLING=[256, 64, 16], MAX_SAMPLES=[64, 256, 1024], KNN=[1, 20, 20], KNN_DILATION=1)
    small = util.build_network(DEV, cfg=cfg)
    assert not small.fusable()
    x = util.synthetic_batch(2, 1024, start=320).to(DEV)
    before = L.lib().pab_num_launches()
    with torch.no_grad():
        d, f, c = small(x)
    assert L.lib().pab_num_launches() > before and torch.isfinite(d).all() and d.shape == (2, 256)


def test_engine_refolds_after_in_place_weight_updates_in_eval_mode(net):
    """optimizer.step() / in-place edits / a child's load_state_dict while the model stays in eval() must not leave
    a stale folded copy behind."""
    other = util.build_network(DEV, seed=777)
    x = util.synthetic_batch(1, 4096, 9).to(DEV)
    with torch.no_grad():
        a, _, _ = other(x)
        other.aggregation.afa.fc.weight.mul_(0.5)                       # in place, eval mode
        b, _, _ = other(x)
        other.backbone.load_state_dict(net.backbone.state_dict())       # a CHILD's load_state_dict
        other.aggregation.load_state_dict(net.aggregation.state_dict())
        other.decoder.load_state_dict(net.decoder.state_dict())
        c, _, _ = other(x)
        want, _, _ = net(x)
    assert not torch.equal(a, b) and torch.equal(c, want)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_eval_forward_matches_single_gpu(net):
    """the reference wraps the model in nn.DataParallel when several GPUs are visible
    (train_place_recognition.py:546-548); replicas have no parameters() and must not touch the master's fused engine."""
    x = util.synthetic_batch(4, 4096, start=330).to(DEV)
    dp = torch.nn.DataParallel(net, device_ids=[0, 1])
    with torch.no_grad():
        want, _, cw = net(x)
        got, _, cg = dp(x)
    assert (got - want).abs().max().item() < TOL and torch.equal(cg[0].cpu(), cw[0].cpu())
