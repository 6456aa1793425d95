#!/usr/bin/env python
"""Per-stage GPU time of the fused PPT-Net engine (events after every C-ABI launch), batch 64 x 4096."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from patchaugnet_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = util.build_pptnet("cuda")
x = torch.cat([util.synthetic_batch(16, 4096, 0)] * (B // 16)).cuda()
with torch.no_grad():
    for _ in range(2): net(x)
torch.cuda.synchronize()
orig = L.check
events = []
def chk(rc, what="call"):
    orig(rc, what)
    e = torch.cuda.Event(enable_timing=True); e.record(); events.append((what, e))
L.check = chk
tot = {}
with torch.no_grad():
    for rep in range(3):
        events.clear()
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
        net(x, return_feat=False)
        torch.cuda.synchronize()
        prev = e0
        counts = {}
        for what, e in events:
            i = counts.get(what, 0); counts[what] = i + 1
            tot.setdefault(f"{what}{i}", []).append(prev.elapsed_time(e)); prev = e
L.check = orig
res = {k: round(sorted(v)[len(v)//2], 4) for k, v in tot.items()}
print(json.dumps(dict(batch=B, total_ms=round(sum(res.values()), 3), stages=dict(sorted(res.items(), key=lambda kv: -kv[1]))), indent=1))
