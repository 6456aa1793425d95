#!/usr/bin/env python
"""Timeline of CTA 0 of the fused-MLP tensor-core kernel for one stage (debugging aid, see pab_tune_tc_trace)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from patchaugnet_b200 import _lib as L

stage = sys.argv[1] if len(sys.argv) > 1 else "fp0"
dev = torch.device("cuda", 0)
net = util.build_network(dev)
eng = net.engine()
x = util.synthetic_batch(32, 4096).to(dev)
with torch.no_grad():
    eng(x); eng(x)
torch.cuda.synchronize()
buf = torch.zeros(8 * 4 * 8 + 256 + 2 * 148, dtype=torch.int64, device=dev)
# run the forward with tracing on: every TC launch overwrites the buffer, so snapshot right after the wanted stage
lib = L.lib()
if os.environ.get("PAB_TC_TUNE"):
    lib.pab_tune_tensor_core(int(os.environ["PAB_TC_TUNE"]))
orig_run = eng._runner
snap = {}
def runner():
    run = orig_run()
    def wrapped(st, fn):
        if st == stage:
            lib.pab_tune_tc_trace(L.ptr(buf)); buf.zero_()
        run(st, fn)
        if st == stage:
            torch.cuda.synchronize(); snap["t"] = buf.clone(); lib.pab_tune_tc_trace(L.ptr(None))
    return wrapped
eng._runner = runner
with torch.no_grad():
    eng(x)
cta = snap["t"].cpu()[512:].view(148, 2)
fine = snap["t"].cpu()[256:512]
t = snap["t"].cpu()[:256].view(8, 4, 8)
t0 = int(t[t > 0].min())
names = ["mma_start", "mma_end", "acc_seen", "epi_done", "load_start", "staged", "stored", "acc0_seen"]
for tile in range(6):
    for l in range(4):
        row = t[tile, l]
        if (row > 0).any():
            print(f"tile {tile} layer {l}: " + "  ".join(f"{names[e]}={int(row[e]) - t0:>7d}" for e in range(8) if row[e] > 0))

st, en = cta[:, 0], cta[:, 1]
ok = st > 0
base = int(st[ok].min())
dur = (en[ok] - st[ok]).float() / 1000.0
print(f"per-CTA wall time (us): min {dur.min():.1f} median {dur.median():.1f} max {dur.max():.1f}; start skew max {(int(st[ok].max()) - base) / 1000.0:.1f} us; "
      f"kernel span {(int(en[ok].max()) - base) / 1000.0:.1f} us")
print("slowest CTAs:", [(int(i), round(float(d), 1)) for d, i in sorted(zip(dur.tolist(), torch.nonzero(ok).flatten().tolist()))[-6:]])
