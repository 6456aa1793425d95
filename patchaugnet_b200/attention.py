"""Host side of the fused PPT-Net self-attention layer (csrc/attention.cu, C ABI ``pab_sa_layer_forward``).

``sa_layer_forward(layer, x)`` evaluates ``pptnet.SA_Layer`` in eval mode: it folds the layer's parameters into the two
point-wise layers the kernel expects — the grouped, tied q/k projection expanded to a dense block-diagonal matrix next to
the v projection, and trans_conv with after_norm folded — and runs the kernels on PyTorch's current stream.
"""
import torch

from . import _lib as L


def _fold(layer, device):
    C = layer.v_conv.weight.shape[0]
    gp = layer.gp
    cg = C // gp
    wk = layer.k_conv.weight.detach().float()[:, :, 0]                     # (C, C/gp): row o uses the inputs of its group
    wq_dense = torch.zeros(C, C, device=wk.device)
    for g in range(gp):
        wq_dense[g * cg:(g + 1) * cg, g * cg:(g + 1) * cg] = wk[g * cg:(g + 1) * cg]
    wv = layer.v_conv.weight.detach().float()[:, :, 0]
    q_wt = wq_dense.t().contiguous().to(device)                                         # (C_in, C_out)
    q_shift = torch.zeros(C, device=device)
    v_wt = wv.t().contiguous().to(device)
    v_shift = layer.v_conv.bias.detach().float().contiguous().to(device)
    bn = layer.after_norm
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    wt = layer.trans_conv.weight.detach().float()[:, :, 0]
    tr_wt = (wt * scale[:, None]).t().contiguous().to(device)                           # (C, C)
    tr_shift = (scale * (layer.trans_conv.bias.detach().float() - bn.running_mean.detach().float())
                + bn.bias.detach().float()).contiguous().to(device)
    # the same three weights as (c_out, c_in) bf16 hi/lo planes: the projections then run on the tensor cores (csrc/pw_tc.cu)
    def planes(w_out_in):
        w = w_out_in.contiguous().to(device)
        hi = w.to(torch.bfloat16)
        return hi.contiguous(), (w - hi.float()).to(torch.bfloat16).contiguous()
    # q and v as the two halves of ONE (2C, C) weight with one (2C) shift: the C side then runs both projections as a single product
    qvh, qvl = planes(torch.cat([wq_dense.to(wv.device), wv], 0))
    qv_shift = torch.cat([q_shift, v_shift]).contiguous()
    q_shift, v_shift = qv_shift[:C], qv_shift[C:]
    th, tl = planes(wt * scale[:, None])
    arr = (L.PabLayer * 3)()
    arr[0] = L.PabLayer(q_wt.data_ptr(), q_shift.data_ptr(), C, C, C, 0, qvh[:C].data_ptr(), qvl[:C].data_ptr(), 0, C)
    arr[1] = L.PabLayer(v_wt.data_ptr(), v_shift.data_ptr(), C, C, C, 0, qvh[C:].data_ptr(), qvl[C:].data_ptr(), 0, C)
    arr[2] = L.PabLayer(tr_wt.data_ptr(), tr_shift.data_ptr(), C, C, C, 1, th.data_ptr(), tl.data_ptr(), 0, C)
    return dict(arr=arr, keep=(q_wt, qv_shift, v_wt, tr_wt, tr_shift, qvh, qvl, th, tl), C=C)


def _versions(layer):
    return tuple(p._version for p in layer.parameters()) + tuple(b._version for b in layer.buffers())


def sa_layer_forward(layer, x, precision=2):
    """x (B, C, N) float32 CUDA -> (B, C, N), eval-mode SA_Layer (pptnet.py:261-282).
    precision: 2 tensor cores with bf16 hi/lo operands (fp32 contract), 1 tensor cores with plain bf16, 0 fp32 SIMT."""
    L.require_cuda(x)
    B, C, N = x.shape
    cache = getattr(layer, "_pab_fold", None)
    ver = _versions(layer)
    if cache is None or cache["ver"] != ver or cache["device"] != x.device:
        cache = _fold(layer, x.device)
        cache["ver"], cache["device"] = ver, x.device
        layer._pab_fold = cache
    xp = x.transpose(1, 2).contiguous().float()                                         # point-major (B, N, C)
    out = torch.empty_like(xp)
    ws = torch.empty(L.lib().pab_sa_layer_workspace_bytes(B, N, C), dtype=torch.uint8, device=x.device)
    arr = cache["arr"]
    L.check(L.lib().pab_sa_layer_forward_p(B, N, C, L.ptr(xp), arr, C_ptr_offset(arr, 1), C_ptr_offset(arr, 2), L.ptr(out),
                                           L.ptr(ws), int(precision), L.stream_ptr()), "sa_layer_forward")
    return out.transpose(1, 2).contiguous()


def C_ptr_offset(arr, i):
    import ctypes
    return ctypes.cast(ctypes.byref(arr, i * ctypes.sizeof(L.PabLayer)), ctypes.POINTER(L.PabLayer))
