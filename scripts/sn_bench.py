"""SA0 (6 -> 32 -> 32 -> 64, k = 20, 32 clouds x 1024 centres) alone: sa_narrow_tc.cu at 1..3 CTAs per SM against mlp_tc.cu's
pre-layer mode.  CUDA events over 50 launches after 5 warm-ups."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import util
from patchaugnet_b200 import _lib as L, pt_util
from patchaugnet_b200.engine import _Layers

DEV = "cuda"
B, n, m, k, c = int(os.environ.get("SN_B", 32)), 4096, 1024, 20, 3
g = torch.Generator().manual_seed(0)
mlp = pt_util.SharedMLP([6, 32, 32, 64], bn=True)
mlp.load_state_dict(util.fill_state_dict(mlp.state_dict(), 1))
mlp = mlp.to(DEV).eval()
layers = _Layers(mlp, DEV, extra_first=3)
xyz = (torch.rand(B, n, 3, generator=g) * 2 - 1).to(DEV)
cidx = torch.stack([torch.randperm(n, generator=g)[:m] for _ in range(B)]).int().to(DEV)
nbr = torch.randint(0, n, (B, m, k), generator=g).int().to(DEV)
out = torch.empty(B, m, 64, device=DEV)


def run():
    L.check(L.lib().pab_sa_module_forward(B, n, m, k, k, c, L.ptr(xyz), L.ptr(xyz), L.ptr(cidx), L.ptr(nbr), layers.arr, layers.n,
                                          L.ptr(out), L.ptr(None), L.stream_ptr()), "sa")


for name, en, per in (("mlp_tc pre-layer mode", 0, 3), ("narrow x1", 1, 1), ("narrow x2", 1, 2), ("narrow x3", 1, 3)):
    L.lib().pab_tune_sa_narrow(en, per)
    for _ in range(5):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:24s} {e0.elapsed_time(e1) / 50 * 1e3:8.1f} us")
L.lib().pab_tune_sa_narrow(1, 0)

# timeline of CTA 0's first steps (clock64 of thread 0): start, pre-layer stored, barrier, MMA1 issued, MMA1 seen, MMA2 issued,
# MMA2 seen, step end
for per, dbg in ((1, 0), (3, 0)):
    L.lib().pab_tune_sa_narrow(1, per)
    L.lib().pab_tune_sa_narrow_dbg(dbg)
    tr = torch.zeros(16 * 8, dtype=torch.int64, device=DEV)
    L.lib().pab_tune_sa_narrow_trace(L.ptr(tr))
    run()
    torch.cuda.synchronize()
    L.lib().pab_tune_sa_narrow_trace(None)
    t = tr.view(16, 8).cpu()
    print(f"-- {per} CTA(s) per SM, dbg {dbg}: cycles since step start [pre stored, barrier, mma1 issued, mma1 seen, mma2 issued, mma2 seen, end], step period")
    for i in range(2, 8):
        row = (t[i] - t[i, 0]).tolist()
        print("   ", row[1:], int(t[i + 1, 0] - t[i, 0]))
L.lib().pab_tune_sa_narrow(1, 0)
L.lib().pab_tune_sa_narrow_dbg(0)
