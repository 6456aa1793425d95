#!/usr/bin/env python
"""PPT-Net: descriptor cosine vs the fp32 golden for every combination of per-part bf16 / hi-lo arithmetic."""
import itertools, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
g = np.load(os.path.join(util.GOLDEN, "pptnet_ref_forward.npz"))
gold = torch.from_numpy(g["desc"]).cuda()
net = util.build_pptnet("cuda")
x = util.golden_batch("pptnet").cuda()
x64 = torch.cat([util.synthetic_batch(16, 4096, 0)] * 4).cuda()
eng = net.engine()
def t_ms():
    for _ in range(2): eng(x64, return_feat=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): eng(x64, return_feat=False)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
out = []
with torch.no_grad():
    for sa, fp, vl, at in itertools.product(("f32", "bf16"), ("f32", "bf16"), ("f32", "bf16"), (2, 1)):
        eng.sa_precision, eng.fp_precision, eng.vlad_precision, eng.attention_precision = sa, fp, vl, at
        eng.refold()
        d = eng(x, return_feat=False)
        cos = torch.nn.functional.cosine_similarity(d, gold).min().item()
        err = (d - gold).abs().max().item()
        out.append(dict(sa=sa, fp=fp, vlad=vl, attn=at, min_cos=round(cos, 6), max_abs=round(err, 6), ms_b64=round(t_ms(), 3)))
        print(json.dumps(out[-1]), flush=True)
