#!/usr/bin/env python
"""torch.profiler breakdown of one PatchAugNet training step (16 anchors x 18 clouds) — where the 300 ms go."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from patchaugnet_b200 import training
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
anchors = int(sys.argv[1]) if len(sys.argv) > 1 else 16
net = util.build_network(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=5e-4)
step = training.TrainStep(net, opt, n_anchors=anchors)
feed = (torch.rand(anchors * 18, 1, 4096, 3, device=dev) * 2 - 1) * 0.57
for _ in range(3): step(feed)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(feed)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): step(feed)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"3 steps: host enqueue {1e3*(t1-t0)/3:.1f} ms/step, total {1e3*(t2-t0)/3:.1f} ms/step")
