"""GPU tier: the training path (BASELINE.json configs[4]) — deterministic backward kernels against the atomic kernels and the
C oracle, and the restated training step (run_model + quadruplet + patch chamfer + Adam) against the same step computed
with the reference's own Python over the reference's own kernels."""
import numpy as np
import pytest
import torch

import util
from oracle import ops
from patchaugnet_b200 import pointops, pointops_cuda, training

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _g(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_deterministic_backward_matches_oracle_and_is_bit_reproducible():
    rng = np.random.default_rng(3)
    b, c, n, m, k = 3, 19, 700, 160, 20
    idx = rng.integers(0, n, (b, m, k)).astype(np.int32)
    idx[0, :, :5] = 7                                             # a heavily shared target: long list, many addends
    g_group = rng.normal(size=(b, c, m, k)).astype(np.float32)
    g_gather = rng.normal(size=(b, c, m)).astype(np.float32)
    idx3 = rng.integers(0, m, (b, n, 3)).astype(np.int32)
    w3 = rng.random((b, n, 3)).astype(np.float32)
    g_interp = rng.normal(size=(b, c, n)).astype(np.float32)

    def run():
        out = []
        g = torch.zeros(b, c, n, device=DEV)
        pointops_cuda.grouping_backward_cuda(b, c, n, m, k, _g(g_group), _g(idx), g)
        out.append(g)
        g = torch.zeros(b, c, n, device=DEV)
        pointops_cuda.gathering_backward_cuda(b, c, n, m, _g(g_gather), _g(idx[:, :, 0].copy()), g)
        out.append(g)
        g = torch.zeros(b, c, m, device=DEV)
        pointops_cuda.interpolation_backward_cuda(b, c, n, m, _g(g_interp), _g(idx3), _g(w3), g)
        out.append(g)
        return out
    assert pointops_cuda.DETERMINISTIC_BACKWARD
    first, second = run(), run()
    for x, y in zip(first, second):
        assert torch.equal(x, y)                                  # bit-identical from run to run
    pointops_cuda.DETERMINISTIC_BACKWARD = False
    try:
        atomic = run()
    finally:
        pointops_cuda.DETERMINISTIC_BACKWARD = True
    want = [ops.grouping_backward(g_group, idx, n), ops.gathering_backward(g_gather, idx[:, :, 0].copy(), n),
            ops.interpolation_backward(g_interp, idx3, w3, m)]
    for det, at, w in zip(first, atomic, want):
        scale = np.abs(w).max()
        assert np.abs(det.cpu().numpy() - w).max() < 1e-5 * scale
        assert (det - at).abs().max().item() < 1e-5 * scale
    # ascending-entry order is exactly the oracle's sequential order for the unweighted scatters
    assert np.array_equal(first[0].cpu().numpy(), want[0]) and np.array_equal(first[1].cpu().numpy(), want[1])


def _small_cfg():
    return dict(util.PATCHAUGNET_CFG, SAMPLING=[256, 64, 16], MAX_SAMPLES=[64, 256, 1024])


def test_training_step_runs_decreases_loss_and_is_reproducible():
    """Two anchors x 18 clouds x 1024 points: the step is deterministic up to cuDNN (scatter kernels are), produces finite
    gradients for every trained parameter, and a few steps reduce the loss."""
    torch.manual_seed(0)
    feed = torch.cat([util.place_batch(range(500, 518), 0, 1024), util.place_batch(range(600, 618), 0, 1024)]).to(DEV)
    losses_run = []
    for rep in range(2):
        torch.manual_seed(1)
        net = util.build_network(DEV, cfg=_small_cfg()).train()
        opt = torch.optim.Adam(net.parameters(), lr=5e-4)
        step = training.TrainStep(net, opt, n_anchors=2)
        assert len(step.nn_dict) == 4 and (18, 20) in step.nn_dict
        seq = []
        for it in range(4):
            torch.manual_seed(100 + it)                            # the randperm of QueryAndGroup_Edge
            loss, terms = step(feed)
            seq.append(loss.item())
            assert set(terms) == {"place_recognition", "patch_recon_a2a"} and all(torch.isfinite(v) for v in terms.values())
        losses_run.append(seq)
        grads = [p.grad for p in net.parameters() if p.grad is not None]
        assert len(grads) > 60 and all(torch.isfinite(g).all() for g in grads)
    assert losses_run[0][-1] < losses_run[0][0]
    assert np.allclose(losses_run[0], losses_run[1], rtol=2e-2)      # cuDNN convolution backward is not bit-reproducible; the step is chaotic (see below)


def test_training_step_matches_the_reference_python_on_the_reference_kernels():
    """Same weights, same tuple batch, same randperm draws: loss and gradients of this repo's step (mirror modules +
    deterministic kernels) against the reference's Network / pointops.py / losses running on the stock kernels."""
    from oracle import refgpu, refpy
    if not (refpy.available() and refgpu.available() and refgpu.torch_kernels_available()):
        pytest.skip("oracle/_ref not built")
    cfg = _small_cfg()
    feed = util.place_batch(range(700, 718), 0, 1024).to(DEV)
    ours = util.build_network(DEV, cfg=cfg).train()
    sd = {k: v.clone() for k, v in ours.state_dict().items()}
    ref = refpy.use_backend("stock")
    ref_net = ref.patch_aug_net.Network(param=dict(ref.cfg_patchaugnet, SAMPLING=cfg["SAMPLING"], MAX_SAMPLES=cfg["MAX_SAMPLES"]),
                                        use_a2a_recon=True, use_l2_norm=True)
    ref_net.load_state_dict(sd)
    ref_net = ref_net.to(DEV).train()
    nn_dict = training.make_nn_dict(1)

    from patchaugnet_b200 import losses
    g = torch.Generator(device=DEV).manual_seed(9)
    R = torch.randn(18, 256, device=DEV, generator=g)

    def smooth_loss(desc, recon):
        # The shipped loss takes a max over positives and over negatives (quadruplet, lazy): which tuple member is "hardest" is
        # a discrete choice that flips under 1e-6 perturbations and makes gradients jump, so gradient parity is checked under a
        # smooth surrogate — every descriptor element weighted by a fixed random matrix + the patch chamfer term — and the
        # shipped loss by value
        return (desc * R).sum() + 0.25 * losses.patch_chamfer_loss(recon["origin_patches"], recon["reconstructed_patches"])

    torch.manual_seed(42)
    x1 = feed.clone().requires_grad_(True)
    desc, recon = ours(x1, nn_dict, return_feat=False)
    loss_shipped, _ = training.assemble_loss(desc.detach(), None, 1)
    loss = smooth_loss(desc, recon)
    loss.backward()

    torch.manual_seed(42)
    x2 = feed.clone().requires_grad_(True)
    (rdesc, rrecon) = ref_net(x2, nn_dict, return_feat=False)
    rloss_shipped, _ = training.assemble_loss(rdesc.detach(), None, 1)
    rloss = smooth_loss(rdesc, rrecon)
    rloss.backward()
    # Fused BatchNorm + ReLU kernels (csrc/bn_train.cu) vs the reference's cuDNN BatchNorm: they differ by 2e-6 per layer in the
    # forward — and are the closer of the two to a float64 evaluation (test_fused_train_bn_relu_matches_torch).  This step is
    # chaotic at that scale: perturbing the WEIGHTS of the cuDNN path by 2e-6 relative moves the loss by 4e-3 and individual
    # gradient tensors by 15-40 % of their maximum (18-cloud batch statistics, max-pool / ReLU / chamfer arg-min flips;
    # scripts/grad_noise.py, profiles/r02_results.md), so element-wise gradient parity is only meaningful with the same
    # BatchNorm arithmetic on both sides (below); here: the forward, the loss values and the overall gradient direction.
    assert (desc - rdesc).abs().max().item() < 5e-3
    assert abs(loss_shipped.item() - rloss_shipped.item()) < 1e-2 * max(1.0, abs(rloss_shipped.item()))
    assert abs(loss.item() - rloss.item()) < 1e-2 * max(1.0, abs(rloss.item()))
    gp = dict(ours.named_parameters())
    gmax = max(p.grad.abs().max().item() for p in ref_net.parameters() if p.grad is not None)
    mine = torch.cat([gp[n].grad.flatten() for n, p in ref_net.named_parameters() if p.grad is not None])
    theirs = torch.cat([p.grad.flatten() for n, p in ref_net.named_parameters() if p.grad is not None])
    assert torch.isfinite(mine).all()
    assert torch.nn.functional.cosine_similarity(mine, theirs, dim=0).item() > 0.97
    for name, p in ref_net.named_parameters():
        if p.grad is None:
            assert gp[name].grad is None or gp[name].grad.abs().max().item() == 0
    # and with the same BatchNorm arithmetic on both sides (cuDNN) the agreement is at the reference's own noise level
    from patchaugnet_b200 import pt_util
    pt_util.FUSED_TRAIN_BN_RELU = False
    try:
        ours.load_state_dict(sd)
        ours.zero_grad(set_to_none=True)
        torch.manual_seed(42)
        x3 = feed.clone().requires_grad_(True)
        desc3, recon3 = ours(x3, nn_dict, return_feat=False)
        smooth_loss(desc3, recon3).backward()
    finally:
        pt_util.FUSED_TRAIN_BN_RELU = True
    assert torch.equal(desc3, rdesc)
    for name, p in ref_net.named_parameters():
        if p.grad is not None:
            assert (gp[name].grad - p.grad).abs().max().item() < 5e-3 * max(p.grad.abs().max().item(), 1e-4 * gmax), name


@pytest.mark.parametrize("shape", [(6, 64, 256, 20), (5, 32, 1000, 1), (3, 7, 13, 5), (4, 256, 333), (18, 256, 64, 1), (18, 512, 16, 20),
                                   (18, 256, 1024, 1), (18, 32, 256, 20), (2, 256, 16, 1), (1, 64, 4, 1)])
def test_fused_train_bn_relu_matches_torch(shape):
    """csrc/bn_train.cu against nn.BatchNorm (train mode) + ReLU: outputs, running statistics, and all three gradients."""
    from patchaugnet_b200 import pt_util
    torch.manual_seed(sum(shape))
    C = shape[1]
    x = (torch.randn(*shape, device=DEV) * 2 + 0.7)
    dy = torch.randn(*shape, device=DEV)
    bn_cls = torch.nn.BatchNorm2d if len(shape) == 4 else torch.nn.BatchNorm1d
    ref = bn_cls(C).to(DEV).train()
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5); ref.bias.normal_(0, 0.3); ref.running_mean.normal_(); ref.running_var.uniform_(0.5, 2)
    w, b = ref.weight.detach().clone().requires_grad_(True), ref.bias.detach().clone().requires_grad_(True)
    rm, rv = ref.running_mean.clone(), ref.running_var.clone()
    x1 = x.clone().requires_grad_(True)
    y1 = torch.relu(ref(x1))
    y1.backward(dy)
    x2 = x.clone().requires_grad_(True)
    y2 = pt_util._BnReluTrain.apply(x2, w, b, rm, rv, ref.momentum, ref.eps)
    y2.backward(dy)
    assert (y1 - y2).abs().max().item() < 2e-5 * max(1.0, y1.abs().max().item())
    # float64 evaluation as the arbiter: the fused kernels (double-precision reductions) are at least as close to it as cuDNN
    ref64 = bn_cls(C).to(DEV).double().train()
    with torch.no_grad():
        ref64.weight.copy_(w.double()); ref64.bias.copy_(b.double())
    x64 = x.double().requires_grad_(True)
    y64 = torch.relu(ref64(x64))
    y64.backward(dy.double())
    e_cudnn, e_ours = (y1.double() - y64).abs().max().item(), (y2.double() - y64).abs().max().item()
    g_cudnn, g_ours = (x1.grad.double() - x64.grad).abs().max().item(), (x2.grad.double() - x64.grad).abs().max().item()
    assert e_ours <= 2 * e_cudnn + 1e-6 and g_ours <= 2 * g_cudnn + 1e-6 * max(1.0, x64.grad.abs().max().item()), (e_ours, e_cudnn, g_ours, g_cudnn)
    assert torch.allclose(rm, ref.running_mean, atol=1e-6, rtol=1e-5) and torch.allclose(rv, ref.running_var, atol=1e-6, rtol=1e-5)
    for a, bb in ((x1.grad, x2.grad), (ref.weight.grad, w.grad), (ref.bias.grad, b.grad)):
        assert (a - bb).abs().max().item() < 1e-4 * max(1.0, a.abs().max().item())
    # deterministic
    x3 = x.clone().requires_grad_(True)
    rm3, rv3 = ref.running_mean.clone(), ref.running_var.clone()
    y3 = pt_util._BnReluTrain.apply(x3, w, b, rm3, rv3, ref.momentum, ref.eps)
    y3.backward(dy)
    assert torch.equal(y3, y2) and torch.equal(x3.grad, x2.grad)


def test_graphed_training_step_matches_the_eager_step():
    """training.GraphedTrainStep: the whole step replayed as one CUDA graph must update the weights like the eager step."""
    import copy
    cfg = _small_cfg()
    feed = util.place_batch(range(800, 818), 0, 1024).to(DEV)
    torch.manual_seed(3)
    base = util.build_network(DEV, cfg=cfg).train()
    nets = [copy.deepcopy(base) for _ in range(2)]
    eager = training.TrainStep(nets[0], torch.optim.Adam(nets[0].parameters(), lr=1e-4, capturable=True), n_anchors=1)
    graphed = training.GraphedTrainStep(nets[1], torch.optim.Adam(nets[1].parameters(), lr=1e-4, capturable=True), n_anchors=1, warmup=2)
    # identical schedules: 2 warm-up steps + capture step (3 optimiser steps) then 3 replays; the eager twin takes 6 steps
    torch.manual_seed(5)
    l_g = [graphed(feed)[0].item()]
    for _ in range(3):
        l_g.append(graphed(feed)[0].item())
    torch.manual_seed(5)
    l_e = [eager(feed)[0].item() for _ in range(6)]
    assert all(np.isfinite(l_g)) and l_g[-1] < l_g[0] + 1e-3
    assert abs(l_g[-1] - l_e[-1]) < 5e-2 * max(1.0, abs(l_e[-1]))
    diffs = [(a - b).abs().max().item() / max(b.abs().max().item(), 1e-6) for a, b in zip(nets[1].parameters(), nets[0].parameters())]
    assert max(diffs) < 5e-2          # same trajectory up to the chaos of the step (different neighbour order, cuDNN noise)
