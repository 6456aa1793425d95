#!/usr/bin/env python
"""FPS timing: full-scan vs pruned sampler, clouds per CTA, uniform vs structured clouds (B=32)."""
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

dev = "cuda"
B = 32
sets = {"uniform": util.synthetic_batch(B, 4096, 0).squeeze(1).to(dev),
        "places": torch.from_numpy(np.stack([util.place_visit(i, 0) for i in range(B)])).to(dev)}
out = {}
for name, x in sets.items():
    for (n, m) in [(4096, 1024), (2048, 512)]:
        xx = x[:, :n].contiguous()
        for pruned in (0, 1):
            for cpc in ((1,) if not pruned else (1, 2, 3)):
                L.lib().pab_tune_fps_pruned(pruned); L.lib().pab_tune_fps_clouds_per_cta(cpc)
                out[f"{name}_n{n}_m{m}_pruned{pruned}_cpc{cpc}"] = round(timed(lambda: pointops.furthestsampling(xx, m)), 4)
L.lib().pab_tune_fps_pruned(1); L.lib().pab_tune_fps_clouds_per_cta(1)
print(json.dumps(out, indent=1))
