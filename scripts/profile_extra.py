#!/usr/bin/env python
"""Driver for the ncu capture of the kernels bench.py's step does not launch: PPT-Net forward (tensor-core attention), retrieval
top-k (250 x 10k, k = 101), one training step at 4 anchors (fused BatchNorm+ReLU, deterministic scatter, chamfer)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from patchaugnet_b200 import retrieval, training
dev = torch.device("cuda", 0)
ppt = util.build_pptnet(dev)
x = torch.cat([util.synthetic_batch(16, 4096, 0)] * 4).to(dev)
with torch.no_grad():
    for _ in range(2): ppt(x, return_feat=False)
db = torch.nn.functional.normalize(torch.randn(10000, 256, device=dev)); q = torch.nn.functional.normalize(torch.randn(250, 256, device=dev))
for _ in range(2): retrieval.retrieval_topk(db, q, 101)
net = util.build_network(dev).train()
step = training.TrainStep(net, torch.optim.Adam(net.parameters(), lr=5e-4), n_anchors=4)
feed = (torch.rand(4 * 18, 1, 4096, 3, device=dev) * 2 - 1) * 0.57
for _ in range(2): step(feed)
torch.cuda.synchronize()
print("done")
