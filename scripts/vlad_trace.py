#!/usr/bin/env python
"""Timeline of CTA 0 of the NetVLAD tensor-core kernel for one level (debugging aid, see pab_tune_tc_trace)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from patchaugnet_b200 import _lib as L

stage = sys.argv[1] if len(sys.argv) > 1 else "vlad2"
dev = torch.device("cuda", 0)
net = util.build_network(dev)
eng = net.engine()
x = util.synthetic_batch(32, 4096).to(dev)
with torch.no_grad():
    eng(x); eng(x)
torch.cuda.synchronize()
buf = torch.zeros(8 * 4 * 8 + 256 + 2 * 148, dtype=torch.int64, device=dev)
lib = L.lib()
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
t = snap["t"].cpu()[:128].view(16, 8)
t0 = int(t[t > 0].min())
names = ["tile_start", "planes_free", "staged", "logits_seen", "act_staged"]
for tile in range(16):
    row = t[tile]
    if (row > 0).any():
        print(f"tile {tile}: " + "  ".join(f"{names[e]}={int(row[e]) - t0:>7d}" for e in range(5) if row[e] > 0))
