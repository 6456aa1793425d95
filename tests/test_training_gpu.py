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
    assert np.allclose(losses_run[0], losses_run[1], rtol=1e-4)


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

    torch.manual_seed(42)
    x1 = feed.clone().requires_grad_(True)
    desc, recon = ours(x1, nn_dict, return_feat=False)
    loss, terms = training.assemble_loss(desc, recon, 1)
    loss.backward()

    torch.manual_seed(42)
    x2 = feed.clone().requires_grad_(True)
    (rdesc, rrecon) = ref_net(x2, nn_dict, return_feat=False)
    # the reference's chamfer module needs its own compiled extension name (`import chamfer`): use this repo's loss functions on
    # the reference's outputs — the forward / backward of the network is what is compared
    rloss, rterms = training.assemble_loss(rdesc, rrecon, 1)
    rloss.backward()
    assert abs(loss.item() - rloss.item()) < 1e-4 * max(1.0, abs(rloss.item()))
    gp = dict(ours.named_parameters())
    n_checked = 0
    for name, p in ref_net.named_parameters():
        if p.grad is None:
            assert gp[name].grad is None or gp[name].grad.abs().max().item() == 0
            continue
        scale = max(p.grad.abs().max().item(), 1e-6)
        assert (gp[name].grad - p.grad).abs().max().item() < 2e-3 * scale, name
        n_checked += 1
    assert n_checked > 60
    assert (x1.grad - x2.grad).abs().max().item() < 2e-3 * max(x2.grad.abs().max().item(), 1e-6)
