"""AFA head: SIMT / tensor-core kernels against float64, with the intermediate softmax weights (debugging aid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from patchaugnet_b200 import _lib as L
DEV = "cuda"
lib = L.lib()
for b, c, K, c_out in [(3, 64, 8, 32), (1, 192, 3, 64), (3, 64, 8, 256), (3, 256, 8, 32), (5, 256, 84, 256)]:
    g = torch.Generator().manual_seed(b * 1000 + K)
    v = torch.randn(b, c, K, generator=g).to(DEV)
    watt = (torch.randn(c, c, generator=g) / c ** 0.5).to(DEV)
    wfc = (torch.randn(c_out, c * K, generator=g) / (c * K) ** 0.5).to(DEV)
    scale = (torch.rand(c_out, generator=g) + 0.5).to(DEV)
    shift = torch.randn(c_out, generator=g).to(DEV)
    vd = v.double()
    att = torch.einsum("oc,bck->bok", watt.double(), vd).max(dim=1)[0]
    w = torch.softmax(att, dim=1)
    y = torch.relu(vd + vd * w[:, None, :]).reshape(b, c * K)
    raw = y @ wfc.double().t() * scale.double() + shift.double()
    want = raw / raw.norm(dim=1, keepdim=True).clamp_min(1e-12)
    split = lambda t: (t.to(torch.bfloat16).contiguous(), (t - t.to(torch.bfloat16).float()).to(torch.bfloat16).contiguous())
    for tc in (0, 1):
        desc = torch.zeros(b, c_out, device=DEV)
        if tc:
            ws = torch.zeros(lib.pab_afa_tc_workspace_bytes(b, c, K, c_out), dtype=torch.uint8, device=DEV)
            (ah, al), (fh, fl) = split(watt), split(wfc)
            L.check(lib.pab_afa_forward_tc(b, c, K, c_out, L.ptr(v), L.ptr(ah), L.ptr(al), L.ptr(fh), L.ptr(fl), L.ptr(scale), L.ptr(shift),
                                           1, L.ptr(desc), L.ptr(ws), L.stream_ptr()), "tc")
        else:
            ws = torch.zeros(lib.pab_afa_workspace_bytes(b, c, K, c_out), dtype=torch.uint8, device=DEV)
            watt_t, wfc_t = watt.t().contiguous(), wfc.t().contiguous()          # (c_in, c_out') / (C*K, c_out): keep them alive
            L.check(lib.pab_afa_forward(b, c, K, c_out, L.ptr(v), L.ptr(watt_t), L.ptr(wfc_t), L.ptr(scale),
                                        L.ptr(shift), 1, L.ptr(desc), L.ptr(ws), L.stream_ptr()), "simt")
        torch.cuda.synchronize()
        wsm = ws[:b * K * 4].view(torch.float32).view(b, K)
        print((b, c, K, c_out), "tc" if tc else "simt", "desc err %.2e" % (desc.double() - want).abs().max().item(),
              "softmax err %.2e" % (wsm.double() - w).abs().max().item(),
              "norm of desc %.4f" % desc.norm(dim=1).mean().item())
