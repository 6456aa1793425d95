#!/usr/bin/env python
"""FPS timing vs batch size, with and without the exclusive-SM shared-memory pad."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from patchaugnet_b200 import _lib as L, pointops

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

x = torch.cat([util.synthetic_batch(16, 4096, 0)] * 8).squeeze(1).cuda()
out = {}
for excl in (0, 1):
    L.lib().pab_tune_fps_exclusive(excl)
    for B in (1, 16, 32, 64, 128):
        for (n, m) in [(4096, 1024), (1024, 256)]:
            xx = x[:B, :n].contiguous()
            out[f"excl{excl}_B{B}_n{n}_m{m}"] = round(timed(lambda: pointops.furthestsampling(xx, m)), 4)
L.lib().pab_tune_fps_exclusive(1)
print(json.dumps(out, indent=1))
