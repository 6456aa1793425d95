"""GPU tier: the tcgen05 tensor-core SharedMLP kernels (mlp_tc.cu) against the fp32 SIMT kernels (mlp.cu, themselves
checked against the oracle) and against a float64 torch evaluation, through the same C-ABI entry points.
Tolerance: the bf16 hi/lo split carries ~16 mantissa bits per product -> relative error ~1e-5 per layer."""
import numpy as np
import pytest
import torch

import util
from patchaugnet_b200 import _lib as L
from patchaugnet_b200 import pt_util
from patchaugnet_b200.engine import _Layers

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mlp(spec, seed):
    torch.manual_seed(seed)
    m = pt_util.SharedMLP(list(spec), bn=True)
    sd = util.fill_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    return m.to(DEV).eval()


def _run_fp(layers, B, n, m, known, skip, idx, w, tc):
    out = torch.empty(B, n, layers.c_out, device=DEV)
    L.lib().pab_tune_tensor_core(1 if tc else 0)
    try:
        L.check(L.lib().pab_fp_module_forward(B, n, m, known.shape[2], 0 if skip is None else skip.shape[2], L.ptr(known), L.ptr(skip),
                                              L.ptr(idx), L.ptr(w), layers.arr, layers.n, L.ptr(out), L.stream_ptr()), "fp")
        torch.cuda.synchronize()
    finally:
        L.lib().pab_tune_tensor_core(1)
    return out


@pytest.mark.parametrize("spec,c_known,c_skip,n,m", [([259, 256, 256, 256], 256, 3, 1000, 300), ([320, 256, 256], 256, 64, 700, 128),
                                                       ([128, 64, 128], 128, 0, 130, 40),
                                                       ([768, 256, 256], 512, 256, 300, 16),      # FP2: 12 chunks in 3 operand groups
                                                       ([384, 256, 256], 256, 128, 517, 64)])     # PPT-Net FP2: 6 chunks = 4 + 2
def test_fp_module_tensor_core_matches_simt_and_fp64(spec, c_known, c_skip, n, m):
    B = 3
    g = torch.Generator(device="cpu").manual_seed(n)
    mlp = _mlp(spec, n)
    layers = _Layers(mlp, DEV, extra_last=c_skip if c_skip <= 3 else 0)
    assert layers.tensor_core
    known = torch.randn(B, m, c_known, generator=g).to(DEV)
    skip = torch.randn(B, n, c_skip, generator=g).to(DEV) if c_skip else None
    idx = torch.randint(0, m, (B, n, 3), generator=g).int().to(DEV)
    w = torch.rand(B, n, 3, generator=g)
    w = (w / w.sum(2, keepdim=True)).to(DEV)
    simt = _run_fp(layers, B, n, m, known, skip, idx, w, tc=False)
    before = L.lib().pab_num_launches()
    tcore = _run_fp(layers, B, n, m, known, skip, idx, w, tc=True)
    assert L.lib().pab_num_launches() == before + 1
    # float64 reference of the same module: interpolate, concat, conv/bn/relu
    gk = torch.gather(known.double(), 1, idx.long().view(B, n * 3, 1).expand(-1, -1, c_known)).view(B, n, 3, c_known)
    x = (gk * w.double().unsqueeze(-1)).sum(2)
    if skip is not None:
        x = torch.cat([x, skip.double()], 2)
    ref = mlp.double()(x.transpose(1, 2).unsqueeze(-1)).squeeze(-1).transpose(1, 2)
    scale = ref.abs().max().item()
    assert (simt.double() - ref).abs().max().item() < 2e-5 * scale
    assert (tcore.double() - ref).abs().max().item() < 5e-5 * scale
    assert torch.isfinite(tcore).all()


def test_tensor_core_scheduling_options_do_not_change_results():
    """Weight multicast across CTA pairs (bit 2) and dynamic tile scheduling (bit 3) only change WHO computes a tile and how
    the weights reach shared memory: outputs must be bit-identical to the default static single-CTA schedule."""
    B, n, m, c_known, c_skip = 8, 4096, 1024, 256, 3            # 256 tiles > 148 CTAs
    g = torch.Generator(device="cpu").manual_seed(5)
    mlp = _mlp([259, 256, 256, 256], 5)
    layers = _Layers(mlp, DEV, extra_last=3)
    known = torch.randn(B, m, c_known, generator=g).to(DEV)
    skip = torch.randn(B, n, c_skip, generator=g).to(DEV)
    idx = torch.randint(0, m, (B, n, 3), generator=g).int().to(DEV)
    w = torch.rand(B, n, 3, generator=g)
    w = (w / w.sum(2, keepdim=True)).to(DEV)
    outs = {}
    for tune in (1, 5, 9):
        out = torch.empty(B, n, layers.c_out, device=DEV)
        L.lib().pab_tune_tensor_core(tune)
        try:
            L.check(L.lib().pab_fp_module_forward(B, n, m, c_known, c_skip, L.ptr(known), L.ptr(skip), L.ptr(idx), L.ptr(w), layers.arr,
                                                  layers.n, L.ptr(out), L.stream_ptr()), "fp")
            torch.cuda.synchronize()
        finally:
            L.lib().pab_tune_tensor_core(1)
        outs[tune] = out
    assert torch.equal(outs[1], outs[5]) and torch.equal(outs[1], outs[9])


@pytest.mark.parametrize("spec,c,k,n,m", [([259, 256, 256, 512], 256, 20, 128, 16), ([67, 64, 64, 256], 64, 20, 1024, 128),
                                           ([131, 128, 128, 256], 128, 7, 300, 50), ([6, 32, 32, 64], 3, 20, 4096, 1024),
                                           ([6, 32, 64], 3, 9, 500, 77)])
def test_sa_module_tensor_core_matches_simt_and_fp64(spec, c, k, n, m):
    B = 4
    g = torch.Generator(device="cpu").manual_seed(n + k)
    mlp = _mlp(spec, n + 1)
    layers = _Layers(mlp, DEV, extra_first=3)
    assert layers.tensor_core
    xyz = (torch.rand(B, n, 3, generator=g) * 2 - 1).to(DEV)
    feat = torch.randn(B, n, c, generator=g).to(DEV)
    cidx = torch.stack([torch.randperm(n, generator=g)[:m] for _ in range(B)]).int().to(DEV)
    nbr = torch.randint(0, n, (B, m, k), generator=g).int().to(DEV)
    outs = []
    for tc in (False, True):
        out = torch.empty(B, m, layers.c_out, device=DEV)
        L.lib().pab_tune_tensor_core(1 if tc else 0)
        try:
            L.check(L.lib().pab_sa_module_forward(B, n, m, k, k, c, L.ptr(xyz), L.ptr(feat), L.ptr(cidx), L.ptr(nbr), layers.arr,
                                                  layers.n, L.ptr(out), L.ptr(None), L.stream_ptr()), "sa")
            torch.cuda.synchronize()
        finally:
            L.lib().pab_tune_tensor_core(1)
        outs.append(out)
    simt, tcore = outs
    scale = simt.abs().max().item()
    assert torch.isfinite(tcore).all()
    assert (simt - tcore).abs().max().item() < 5e-5 * scale
    # float64 torch evaluation of the reference's own op sequence (QueryAndGroup_Edge pointops.py:559-570 ->
    # SharedMLP pt_util.py:98-151 -> max over K patch_aug_net.py:236), independent of both kernels
    li = nbr.long().view(B, m * k)
    g_xyz = torch.gather(xyz.double(), 1, li[..., None].expand(-1, -1, 3)).view(B, m, k, 3)
    g_feat = torch.gather(feat.double(), 1, li[..., None].expand(-1, -1, c)).view(B, m, k, c)
    ctr_xyz = torch.gather(xyz.double(), 1, cidx.long()[..., None].expand(-1, -1, 3))
    ctr_feat = torch.gather(feat.double(), 1, cidx.long()[..., None].expand(-1, -1, c))
    x = torch.cat([g_xyz - ctr_xyz[:, :, None], g_feat - ctr_feat[:, :, None]], -1).permute(0, 3, 1, 2)   # (B, 3+c, m, k)
    with torch.no_grad():
        want = mlp.double()(x).max(dim=3)[0].permute(0, 2, 1)                                             # (B, m, c_out)
    mlp.float()
    assert (tcore.double() - want).abs().max().item() < 5e-5 * want.abs().max().item()
    assert (simt.double() - want).abs().max().item() < 5e-5 * want.abs().max().item()


@pytest.mark.parametrize("n,K", [(4096, 64), (1024, 16), (128, 4), (300, 64)])
def test_netvlad_tensor_core_matches_simt_and_fp64(n, K):
    """vlad_tc.cu (GEMM2 reads the x planes MN-major) vs vlad.cu vs the reference formula in float64."""
    from patchaugnet_b200.engine import _split_bf16
    B, Cf = 3, 256
    g = torch.Generator(device="cpu").manual_seed(n + K)
    x = torch.randn(B, n, Cf, generator=g).to(DEV)
    wc = (torch.randn(Cf, K, generator=g) / 16).to(DEV)
    shift = (torch.randn(K, generator=g) * 0.1).to(DEV)
    w2 = (torch.randn(Cf, K, generator=g) / 16).to(DEV)
    Kp = (K + 15) // 16 * 16
    wct = torch.zeros(Kp, Cf, device=DEV)
    wct[:K] = wc.t()
    hi, lo = _split_bf16(wct)
    lib = L.lib()
    ws = torch.empty(lib.pab_netvlad_workspace_bytes(B, n, Cf, K), dtype=torch.uint8, device=DEV)
    out_s = torch.empty(B, Cf, K, device=DEV)
    out_t = torch.empty(B, Cf, K, device=DEV)
    L.check(lib.pab_netvlad_forward(B, n, Cf, K, L.ptr(x), L.ptr(wc), L.ptr(shift), L.ptr(w2), L.ptr(out_s), Cf * K, K, L.ptr(ws),
                                    L.stream_ptr()), "vlad simt")
    torch.cuda.synchronize()
    L.check(lib.pab_netvlad_forward_tc(B, n, Cf, K, L.ptr(x), L.ptr(hi), L.ptr(lo), L.ptr(shift), L.ptr(w2), L.ptr(out_t), Cf * K, K,
                                       L.ptr(ws), L.stream_ptr()), "vlad tc")
    torch.cuda.synchronize()
    xd = x.double()
    act = torch.softmax(xd @ wc.double() + shift.double(), dim=-1)
    vlad = (act.transpose(1, 2) @ xd).transpose(1, 2) - act.sum(1, keepdim=True) * w2.double()[None]
    ref = torch.nn.functional.normalize(vlad, dim=1, p=2)
    assert (out_s.double() - ref).abs().max().item() < 2e-5
    assert torch.isfinite(out_t).all()
    assert (out_t.double() - ref).abs().max().item() < 5e-5


@pytest.mark.parametrize("spec,c,k,n,m,B", [([6, 32, 32, 64], 3, 20, 4096, 1024, 32), ([6, 32, 64], 3, 9, 500, 77, 4),
                                             ([6, 64, 32, 64], 3, 32, 700, 129, 3), ([3, 32, 32], 0, 128, 256, 5, 2),
                                             ([8, 32, 64, 64, 32], 5, 1, 300, 300, 2), ([6, 32, 32, 64], 3, 20, 64, 1, 1)])
def test_sa_narrow_kernel_bit_identical_to_warp_specialised_kernel(spec, c, k, n, m, B):
    """sa_narrow_tc.cu (three 128-thread CTAs per SM, weights resident, running max in registers, items from a self-resetting counter) against mlp_tc.cu's
    pre-layer mode: same arithmetic, so the outputs must be bit-identical — for every residency setting and on repeated launches
    (the counter has to come back to zero)."""
    g = torch.Generator(device="cpu").manual_seed(n * 7 + k)
    mlp = _mlp(spec, n + 3)
    layers = _Layers(mlp, DEV, extra_first=3)
    assert layers.tensor_core
    xyz = (torch.rand(B, n, 3, generator=g) * 2 - 1).to(DEV)
    feat = torch.randn(B, n, max(c, 1), generator=g).to(DEV)
    cidx = torch.stack([torch.randperm(n, generator=g)[:m] for _ in range(B)]).int().to(DEV)
    nbr = torch.randint(0, n, (B, m, k), generator=g).int().to(DEV)

    def run(enable, per_sm):
        out = torch.full((B, m, layers.c_out), float("nan"), device=DEV)
        L.lib().pab_tune_sa_narrow(enable, per_sm)
        try:
            L.check(L.lib().pab_sa_module_forward(B, n, m, k, k, c, L.ptr(xyz), L.ptr(feat), L.ptr(cidx), L.ptr(nbr), layers.arr,
                                                  layers.n, L.ptr(out), L.ptr(None), L.stream_ptr()), "sa")
            torch.cuda.synchronize()
        finally:
            L.lib().pab_tune_sa_narrow(1, 0)
        return out

    want = run(0, 3)
    assert torch.isfinite(want).all()
    for per_sm in (3, 3, 1, 2, 3):
        got = run(1, per_sm)
        assert torch.equal(got, want), (per_sm, (got - want).abs().max().item())


@pytest.mark.parametrize("b,c,K,c_out", [(32, 256, 84, 256), (5, 256, 84, 256), (3, 64, 8, 32), (130, 128, 20, 128), (1, 192, 3, 64)])
def test_afa_head_tensor_core_matches_simt_and_fp64(b, c, K, c_out):
    """afa_tc.cu (attention logits and fc as bf16 hi/lo tcgen05 GEMMs) against the fp32 SIMT kernels of vlad.cu and a float64
    evaluation of AdaptiveFeatureAggregator.forward (loupe.py:23-60, eval): conv1d -> max over channels -> softmax over clusters
    -> x + x * w -> ReLU -> fc -> BatchNorm (folded) -> L2 norm."""
    g = torch.Generator(device="cpu").manual_seed(b * 1000 + K)
    v = torch.randn(b, c, K, generator=g).to(DEV)
    watt = (torch.randn(c, c, generator=g) / c ** 0.5).to(DEV)                 # (c_out', c_in)
    wfc = (torch.randn(c_out, c * K, generator=g) / (c * K) ** 0.5).to(DEV)    # (c_out, C*K)
    scale = (torch.rand(c_out, generator=g) + 0.5).to(DEV)
    shift = torch.randn(c_out, generator=g).to(DEV)
    lib = L.lib()
    assert lib.pab_afa_tc_supported(c, K, c_out)

    def split(w):
        hi = w.to(torch.bfloat16)
        return hi.contiguous(), (w - hi.float()).to(torch.bfloat16).contiguous()

    outs = []
    for tc in (False, True):
        desc = torch.full((b, c_out), float("nan"), device=DEV)
        if tc:
            ws = torch.empty(lib.pab_afa_tc_workspace_bytes(b, c, K, c_out), dtype=torch.uint8, device=DEV)
            (ah, al), (fh, fl) = split(watt), split(wfc)
            L.check(lib.pab_afa_forward_tc(b, c, K, c_out, L.ptr(v), L.ptr(ah), L.ptr(al), L.ptr(fh), L.ptr(fl), L.ptr(scale), L.ptr(shift),
                                           1, L.ptr(desc), L.ptr(ws), L.stream_ptr()), "afa_tc")
        else:
            ws = torch.empty(lib.pab_afa_workspace_bytes(b, c, K, c_out), dtype=torch.uint8, device=DEV)
            watt_t, wfc_t = watt.t().contiguous(), wfc.t().contiguous()          # (c_in, c_out') / (C*K, c_out): keep them alive
            L.check(lib.pab_afa_forward(b, c, K, c_out, L.ptr(v), L.ptr(watt_t), L.ptr(wfc_t), L.ptr(scale),
                                        L.ptr(shift), 1, L.ptr(desc), L.ptr(ws), L.stream_ptr()), "afa")
        torch.cuda.synchronize()
        outs.append(desc)
    simt, tcore = outs
    vd = v.double()
    att = torch.einsum("oc,bck->bok", watt.double(), vd).max(dim=1)[0]          # (b, K)
    w = torch.softmax(att, dim=1)[:, None, :]
    y = torch.relu(vd + vd * w).reshape(b, c * K)
    want = y @ wfc.double().t() * scale.double() + shift.double()
    want = want / want.norm(dim=1, keepdim=True).clamp_min(1e-12)
    assert torch.isfinite(tcore).all()
    e_simt, e_tc = (simt.double() - want).abs().max().item(), (tcore.double() - want).abs().max().item()
    assert e_simt < 2e-6 and e_tc < 5e-6, (e_simt, e_tc)


@pytest.mark.parametrize("K,N,rows", [(64, 64, 65536), (128, 128, 1000), (256, 256, 4096), (512, 512, 1024), (64, 32, 77), (192, 64, 130)])
def test_pointwise_layer_tensor_core_matches_simt_and_fp64(K, N, rows):
    """pw_tc.cu (one point-wise layer as a K-chunked bf16 hi/lo tcgen05 GEMM) against the fp32 tile kernel of mlp.cu and float64,
    through pab_pointwise_mlp_forward."""
    g = torch.Generator(device="cpu").manual_seed(K + N)
    mlp = _mlp([K, N], K + 1)
    layers = _Layers(mlp, DEV)
    assert layers.tensor_core
    x = torch.randn(rows, K, generator=g).to(DEV)
    outs = []
    for tc in (0, 1):
        out = torch.full((rows, N), float("nan"), device=DEV)
        L.lib().pab_tune_pointwise_tc(tc)
        try:
            L.check(L.lib().pab_pointwise_mlp_forward(rows, L.ptr(x), layers.arr, layers.n, L.ptr(out), L.stream_ptr()), "pw")
            torch.cuda.synchronize()
        finally:
            L.lib().pab_tune_pointwise_tc(1)
        outs.append(out)
    simt, tcore = outs
    with torch.no_grad():
        want = mlp.double()(x.double().t()[None, :, :, None])[0, :, :, 0].t()
    mlp.float()
    scale = want.abs().max().item()
    assert torch.isfinite(tcore).all()
    assert (simt.double() - want).abs().max().item() < 2e-5 * scale
    assert (tcore.double() - want).abs().max().item() < 5e-5 * scale
    assert not torch.equal(simt, tcore) or K == 0            # the two paths really are different kernels
