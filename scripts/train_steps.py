#!/usr/bin/env python
"""Per-step time of the training step over 12 steps (events), plus SM clock samples: is the step time stable?"""
import os, subprocess, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from patchaugnet_b200 import training, pt_util
dev = torch.device("cuda", 0)
net = util.build_network(dev).train()
step = training.TrainStep(net, torch.optim.Adam(net.parameters(), lr=5e-4), n_anchors=16)
feed = (torch.rand(16 * 18, 1, 4096, 3, device=dev) * 2 - 1) * 0.57
def smi():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active,memory.used", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
times = []
for i in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    step(feed)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    times.append((round(1e3 * (t1 - t0), 1), round(1e3 * (t2 - t0), 1)))
    if i in (2, 6, 11): print("smi:", smi(), flush=True)
print("host/total ms per step:", times)
print("alloc stats: num_alloc_retries", torch.cuda.memory_stats()["num_alloc_retries"], "num_device_alloc", torch.cuda.memory_stats().get("num_device_alloc"), "reserved GB", torch.cuda.memory_reserved() / 2**30)
# the same step replayed as one CUDA graph
import copy
net2 = util.build_network(dev).train()
g = training.GraphedTrainStep(net2, torch.optim.Adam(net2.parameters(), lr=5e-4, capturable=True), n_anchors=16)
t0 = time.perf_counter(); g(feed); torch.cuda.synchronize(); print(f"capture (3 eager warm-up steps + capture): {time.perf_counter() - t0:.2f} s")
times = []
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loss, _ = g(feed)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    times.append((round(1e3 * (t1 - t0), 2), round(1e3 * (t2 - t0), 1)))
print("graphed host/total ms per step:", times, "loss", float(loss))
