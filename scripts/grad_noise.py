#!/usr/bin/env python
"""Gradient agreement of the training step: this repo (fused BN+ReLU on / off) vs the reference Python on stock kernels, and the
reference against itself run twice (its atomics / cuDNN reductions are not deterministic)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from oracle import refpy
from patchaugnet_b200 import pt_util, training
DEV = "cuda"
cfg = dict(util.PATCHAUGNET_CFG, SAMPLING=[256, 64, 16], MAX_SAMPLES=[64, 256, 1024])
feed = util.place_batch(range(700, 718), 0, 1024).to(DEV)
sd = {k: v.clone() for k, v in util.build_network(DEV, cfg=cfg).state_dict().items()}
nn_dict = training.make_nn_dict(1)
def grads(net):
    net.zero_grad(set_to_none=True)
    torch.manual_seed(42)
    x = feed.clone().requires_grad_(True)
    desc, recon = net(x, nn_dict, return_feat=False)
    g = torch.Generator(device=DEV).manual_seed(9)
    R = torch.randn(desc.shape, device=DEV, generator=g)
    from patchaugnet_b200 import losses
    # smooth surrogate (no max over positives / negatives, whose discrete choice makes gradients jump): every descriptor element
    # weighted by a fixed random matrix + the patch chamfer term
    loss = (desc * R).sum() + 0.25 * losses.patch_chamfer_loss(recon["origin_patches"], recon["reconstructed_patches"])
    loss.backward()
    return loss.item(), {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
def cmp(a, b, label):
    worst, name = 0.0, ""
    for n in a:
        if n in b:
            r = (a[n] - b[n]).abs().max().item() / max(b[n].abs().max().item(), 1e-6)
            if r > worst: worst, name = r, n
    print(f"{label}: worst relative (to max) gradient difference {worst:.2e} at {name}")
ours = util.build_network(DEV, cfg=cfg).train(); ours.load_state_dict(sd); ours.train()
ref = refpy.use_backend("stock")
rn = ref.patch_aug_net.Network(param=dict(ref.cfg_patchaugnet, SAMPLING=cfg["SAMPLING"], MAX_SAMPLES=cfg["MAX_SAMPLES"]), use_a2a_recon=True, use_l2_norm=True)
rn.load_state_dict(sd); rn = rn.to(DEV).train()
pt_util.FUSED_TRAIN_BN_RELU = True
l1, g_fused = grads(ours)
ours.load_state_dict(sd)
pt_util.FUSED_TRAIN_BN_RELU = False
l2, g_plain = grads(ours)
pt_util.FUSED_TRAIN_BN_RELU = True
refpy.use_backend("stock")
l3, g_ref1 = grads(rn)
rn.load_state_dict(sd)
l4, g_ref2 = grads(rn)
print("losses", l1, l2, l3, l4)
cmp(g_fused, g_plain, "ours fused-BN vs ours cuDNN-BN")
cmp(g_fused, g_ref1, "ours fused-BN vs reference")
cmp(g_plain, g_ref1, "ours cuDNN-BN vs reference")
cmp(g_ref2, g_ref1, "reference vs reference (second run)")
# conditioning: the cuDNN-BN path against itself with every parameter perturbed by 2e-6 relative (the size of the fused BatchNorm's
# forward difference): how far do the gradients move?
pt_util.FUSED_TRAIN_BN_RELU = False
ours.load_state_dict(sd)
with torch.no_grad():
    g = torch.Generator(device=DEV).manual_seed(5)
    for p in ours.parameters():
        p.mul_(1 + 2e-6 * torch.randn(p.shape, device=DEV, generator=g))
l5, g_pert = grads(ours)
pt_util.FUSED_TRAIN_BN_RELU = True
print("loss perturbed", l5)
cmp(g_pert, g_plain, "ours cuDNN-BN with 2e-6 weight perturbation vs unperturbed")
def cmp_all(a, b, label, k=6):
    rows = sorted(((a[n] - b[n]).abs().max().item() / max(b[n].abs().max().item(), 1e-6), n) for n in a if n in b)
    print(label, [(f"{r:.1e}", n.replace("backbone.", "")) for r, n in rows[-k:]])
cmp_all(g_fused, g_plain, "fused vs cuDNN, worst:")
cmp_all(g_pert, g_plain, "perturbed vs cuDNN, worst:")
