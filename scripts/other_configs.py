#!/usr/bin/env python
"""Timings of the BASELINE.json configurations that are NOT the bench line (bench.py measures configs[1] only):

  configs[2]  PPT-Net pyramid encoder + NetVLAD head, batch 64 x 4096 pts, eval, on the mirrored modules (op by op through
              this repo's kernels: FPS / indexed kNN / grouping, fused SA_Layer attention, SIMT point-wise MLPs) — fp32
              arithmetic (the bf16 of the config line is a throughput hint of the plan, this path keeps fp32 parity)
  configs[4]  PatchAugNet training step (train mode, quadruplet + patch-chamfer loss, autograd through the pointops /
              chamfer Functions), A anchors x 18 clouds x 4096 pts on ONE GPU (the config's 128 anchors over 8 GPUs = 16 per GPU)

Prints one JSON line per configuration.  Reported for completeness; not optimised this round.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def pptnet(dev, batch=64):
    import util
    net = util.build_pptnet(dev)
    x = util.synthetic_batch(batch, 4096, start=0).to(dev)
    with torch.no_grad():
        ms = timed(lambda: net(x))
    print(json.dumps(dict(config="PPT-Net eval, batch %d x 4096, fp32, 1 GPU (BASELINE.json configs[2])" % batch,
                          ms_per_batch=ms, submaps_per_s=batch / (ms * 1e-3))))


def train_step(dev, anchors=2):
    import util
    from patchaugnet_b200 import losses
    from patchaugnet_b200.chamfer_dist import ChamferDistanceL1
    net = util.build_network(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-5)
    per = 18                                     # 1 query + 2 positives + 14 negatives + 1 other negative (train config)
    x = util.synthetic_batch(anchors * per, 4096, start=0).to(dev)
    nn_dict = {(i * per, i * per + 1): [[j, j] for j in range(0, 1024, 64)] for i in range(anchors)}     # overlap pairs for a2a recon
    cd = ChamferDistanceL1()

    def step():
        opt.zero_grad(set_to_none=True)
        out = net(x, nn_dict, return_feat=False)
        desc, recon = out if isinstance(out, tuple) else (out, None)
        d = desc.view(anchors, per, -1)
        q, pos, neg, oth = d[:, :1], d[:, 1:3], d[:, 3:17], d[:, 17:18]
        loss = losses.quadruplet_loss(q, pos, neg, oth, 0.5, 0.2, lazy=True)
        if recon is not None and len(recon["origin_patches"]):
            loss = loss + cd(torch.cat(recon["origin_patches"]), torch.cat(recon["reconstructed_patches"]))
        loss.backward()
        opt.step()
    ms = timed(step, warm=1, reps=3)
    print(json.dumps(dict(config="PatchAugNet training step, %d anchors x 18 clouds x 4096 pts, 1 GPU, quadruplet + patch chamfer "
                                 "(BASELINE.json configs[4] is 16 anchors per GPU)" % anchors,
                          ms_per_step=ms, clouds_per_s=anchors * per / (ms * 1e-3))))


if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    which = sys.argv[1:] or ["pptnet", "train"]
    if "pptnet" in which:
        pptnet(dev)
    if "train" in which:
        train_step(dev)
